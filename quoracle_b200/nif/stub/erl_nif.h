/*
 * STUB erl_nif.h -- NOT the OTP header.  There is no BEAM (and no erl_nif.h) in this image, so the NIF
 * shim cannot be built here; this file declares only the subset of the documented erl_nif API that
 * qconsensus_nif.c uses, so that `gcc -fsyntax-only` can type-check the shim (done in
 * __graft_entry__.build()).  Build the real NIF against OTP's own header.
 */
#ifndef QC_STUB_ERL_NIF_H
#define QC_STUB_ERL_NIF_H
#include <stddef.h>
#include <stdint.h>

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
ErlNifResourceType* enif_open_resource_type(ErlNifEnv*, const char*, const char*, ErlNifResourceDtor*,
                                            ErlNifResourceFlags, ErlNifResourceFlags*);

#define ERL_NIF_INIT(MODNAME, FUNCS, LOAD, RELOAD, UPGRADE, UNLOAD)                                         \
  ErlNifEntry* nif_init(void) {                                                                             \
    static ErlNifEntry entry = {2, 16, #MODNAME, sizeof(FUNCS) / sizeof(*FUNCS), FUNCS, LOAD, RELOAD, UPGRADE, \
                                UNLOAD};                                                                    \
    return &entry;                                                                                          \
  }
#endif
