/*
 * erl_nif.h of the MOCK Erlang runtime used by this repository's tests -- NOT the OTP header.
 *
 * There is no BEAM (and no OTP erl_nif.h) in this image or on the GPU box, so the NIF shim
 * quoracle_b200/nif/qconsensus_nif.c cannot be loaded by a real VM here.  Instead of only type-checking it,
 * tests/mock_erts/ gives it something to RUN against: this header declares the subset of the documented erl_nif
 * API the shim uses (same names, argument order and return conventions as OTP 27's erl_nif.h), and mock_erts.c
 * implements that subset functionally (terms are tagged heap cells owned by an environment, binaries are real
 * byte buffers, resources are reference counted with destructors, mutexes are pthread mutexes).  The -m gpu test
 * tests/test_gpu_nif.py links the unmodified shim against it and drives the NIF entry points on the B200 with
 * packed binaries, exactly as the Elixir collector would.  Build the production NIF against OTP's own header
 * ("Building the NIF"); nothing in the shim depends on this mock.
 */
#ifndef QC_MOCK_ERL_NIF_H
#define QC_MOCK_ERL_NIF_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct enif_environment_t ErlNifEnv;
typedef uintptr_t ERL_NIF_TERM;
typedef struct {
  size_t size;
  unsigned char* data;
  void* ref_bin;
  void* __spare__[2];
} ErlNifBinary;
typedef struct enif_resource_type_t ErlNifResourceType;
typedef void ErlNifResourceDtor(ErlNifEnv*, void*);
typedef enum { ERL_NIF_RT_CREATE = 1, ERL_NIF_RT_TAKEOVER = 2 } ErlNifResourceFlags;
typedef struct ErlNifMutex_ ErlNifMutex;
typedef struct {
  const char* name;
  unsigned arity;
  ERL_NIF_TERM (*fptr)(ErlNifEnv* env, int argc, const ERL_NIF_TERM argv[]);
  unsigned flags;
} ErlNifFunc;
typedef struct {
  int major, minor;
  const char* name;
  int num_of_funcs;
  ErlNifFunc* funcs;
  int (*load)(ErlNifEnv*, void** priv_data, ERL_NIF_TERM load_info);
  int (*reload)(ErlNifEnv*, void** priv_data, ERL_NIF_TERM load_info);
  int (*upgrade)(ErlNifEnv*, void** priv_data, void** old_priv_data, ERL_NIF_TERM load_info);
  void (*unload)(ErlNifEnv*, void* priv_data);
} ErlNifEntry;

#define ERL_NIF_DIRTY_JOB_CPU_BOUND 1
#define ERL_NIF_DIRTY_JOB_IO_BOUND 2

ERL_NIF_TERM enif_make_atom(ErlNifEnv*, const char*);
ERL_NIF_TERM enif_make_tuple2(ErlNifEnv*, ERL_NIF_TERM, ERL_NIF_TERM);
ERL_NIF_TERM enif_make_tuple3(ErlNifEnv*, ERL_NIF_TERM, ERL_NIF_TERM, ERL_NIF_TERM);
ERL_NIF_TERM enif_make_tuple4(ErlNifEnv*, ERL_NIF_TERM, ERL_NIF_TERM, ERL_NIF_TERM, ERL_NIF_TERM);
ERL_NIF_TERM enif_make_badarg(ErlNifEnv*);
ERL_NIF_TERM enif_make_int(ErlNifEnv*, int);
ERL_NIF_TERM enif_make_uint64(ErlNifEnv*, uint64_t);
ERL_NIF_TERM enif_make_resource(ErlNifEnv*, void*);
unsigned char* enif_make_new_binary(ErlNifEnv*, size_t, ERL_NIF_TERM*);
int enif_get_int(ErlNifEnv*, ERL_NIF_TERM, int*);
int enif_inspect_binary(ErlNifEnv*, ERL_NIF_TERM, ErlNifBinary*);
int enif_get_resource(ErlNifEnv*, ERL_NIF_TERM, ErlNifResourceType*, void**);
int enif_is_identical(ERL_NIF_TERM, ERL_NIF_TERM);
void* enif_alloc_resource(ErlNifResourceType*, size_t);
void enif_release_resource(void*);
int enif_keep_resource(void*);
ErlNifResourceType* enif_open_resource_type(ErlNifEnv*, const char*, const char*, ErlNifResourceDtor*,
                                            ErlNifResourceFlags, ErlNifResourceFlags*);
ErlNifMutex* enif_mutex_create(char* name);
void enif_mutex_destroy(ErlNifMutex*);
void enif_mutex_lock(ErlNifMutex*);
void enif_mutex_unlock(ErlNifMutex*);

#define ERL_NIF_INIT(MODNAME, FUNCS, LOAD, RELOAD, UPGRADE, UNLOAD)                                         \
  ErlNifEntry* nif_init(void) {                                                                             \
    static ErlNifEntry entry = {2, 17, #MODNAME, sizeof(FUNCS) / sizeof(*FUNCS), FUNCS, LOAD, RELOAD, UPGRADE, \
                                UNLOAD};                                                                    \
    return &entry;                                                                                          \
  }

#ifdef __cplusplus
}
#endif
#endif
