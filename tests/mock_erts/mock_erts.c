/*
 * mock_erts.c -- a functional stand-in for the part of the Erlang runtime a NIF library talks to (see erl_nif.h in
 * this directory for why it exists).  TEST INFRASTRUCTURE: nothing under quoracle_b200/ links or loads this.
 *
 * Terms are pointers to tagged cells; an environment owns every cell made through it and frees them together (what
 * a process heap + GC does for a real NIF call).  Atoms are interned for the life of the process so that
 * enif_is_identical works across environments.  Resources carry a reference count: enif_alloc_resource starts at
 * 1, a resource TERM holds one reference for as long as its environment lives, enif_release_resource drops one, and
 * the destructor runs when the count reaches zero -- the lifecycle a qc_ctx resource sees in the VM.
 *
 * The mock_* functions are the "VM side": how the test (through ctypes) builds argument terms, calls a NIF by
 * name/arity out of the table nif_init() returns, and takes result terms apart.
 */
#include <pthread.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "erl_nif.h"

enum { K_INT = 1, K_U64, K_ATOM, K_TUPLE, K_BIN, K_RES, K_BADARG };

typedef struct cell {
  int kind;
  long long ival;
  unsigned long long uval;
  const char* atom;
  int arity;
  struct cell** elems;
  unsigned char* data;
  size_t size;
  void* res;
  struct cell* next;   /* env ownership chain */
} cell;

struct enif_environment_t {
  cell* cells;
};

struct enif_resource_type_t {
  char name[64];
  ErlNifResourceDtor* dtor;
};

typedef struct res_hdr {
  ErlNifResourceType* type;
  long refs;
  size_t size;
} res_hdr;

struct ErlNifMutex_ {
  pthread_mutex_t m;
};

static pthread_mutex_t g_lock = PTHREAD_MUTEX_INITIALIZER;
static cell* g_atoms = NULL;         /* interned, never freed */
static long g_live_resources = 0;

static cell* new_cell(ErlNifEnv* env, int kind) {
  cell* c = (cell*)calloc(1, sizeof(cell));
  c->kind = kind;
  c->next = env->cells;
  env->cells = c;
  return c;
}

static void res_unref(void* obj) {
  res_hdr* h = (res_hdr*)obj - 1;
  pthread_mutex_lock(&g_lock);
  long left = --h->refs;
  pthread_mutex_unlock(&g_lock);
  if (left == 0) {
    if (h->type->dtor) h->type->dtor(NULL, obj);
    pthread_mutex_lock(&g_lock);
    --g_live_resources;
    pthread_mutex_unlock(&g_lock);
    free(h);
  }
}

/* ---- erl_nif API ---------------------------------------------------------------------------------------- */
ERL_NIF_TERM enif_make_atom(ErlNifEnv* env, const char* name) {
  (void)env;
  pthread_mutex_lock(&g_lock);
  for (cell* a = g_atoms; a; a = a->next)
    if (strcmp(a->atom, name) == 0) {
      pthread_mutex_unlock(&g_lock);
      return (ERL_NIF_TERM)a;
    }
  cell* a = (cell*)calloc(1, sizeof(cell));
  a->kind = K_ATOM;
  a->atom = strdup(name);
  a->next = g_atoms;
  g_atoms = a;
  pthread_mutex_unlock(&g_lock);
  return (ERL_NIF_TERM)a;
}

static ERL_NIF_TERM make_tuple(ErlNifEnv* env, int n, const ERL_NIF_TERM* e) {
  cell* c = new_cell(env, K_TUPLE);
  c->arity = n;
  c->elems = (cell**)calloc((size_t)n, sizeof(cell*));
  for (int i = 0; i < n; ++i) c->elems[i] = (cell*)e[i];
  return (ERL_NIF_TERM)c;
}
ERL_NIF_TERM enif_make_tuple2(ErlNifEnv* env, ERL_NIF_TERM a, ERL_NIF_TERM b) {
  ERL_NIF_TERM e[2] = {a, b};
  return make_tuple(env, 2, e);
}
ERL_NIF_TERM enif_make_tuple3(ErlNifEnv* env, ERL_NIF_TERM a, ERL_NIF_TERM b, ERL_NIF_TERM c) {
  ERL_NIF_TERM e[3] = {a, b, c};
  return make_tuple(env, 3, e);
}
ERL_NIF_TERM enif_make_tuple4(ErlNifEnv* env, ERL_NIF_TERM a, ERL_NIF_TERM b, ERL_NIF_TERM c, ERL_NIF_TERM d) {
  ERL_NIF_TERM e[4] = {a, b, c, d};
  return make_tuple(env, 4, e);
}
ERL_NIF_TERM enif_make_badarg(ErlNifEnv* env) { return (ERL_NIF_TERM)new_cell(env, K_BADARG); }
ERL_NIF_TERM enif_make_int(ErlNifEnv* env, int v) {
  cell* c = new_cell(env, K_INT);
  c->ival = v;
  return (ERL_NIF_TERM)c;
}
ERL_NIF_TERM enif_make_uint64(ErlNifEnv* env, uint64_t v) {
  cell* c = new_cell(env, K_U64);
  c->uval = v;
  return (ERL_NIF_TERM)c;
}
unsigned char* enif_make_new_binary(ErlNifEnv* env, size_t size, ERL_NIF_TERM* out) {
  cell* c = new_cell(env, K_BIN);
  c->size = size;
  c->data = (unsigned char*)malloc(size ? size : 1);
  *out = (ERL_NIF_TERM)c;
  return c->data;
}
int enif_get_int(ErlNifEnv* env, ERL_NIF_TERM t, int* out) {
  (void)env;
  cell* c = (cell*)t;
  if (c == NULL || c->kind != K_INT || c->ival < -2147483647LL - 1 || c->ival > 2147483647LL) return 0;
  *out = (int)c->ival;
  return 1;
}
int enif_inspect_binary(ErlNifEnv* env, ERL_NIF_TERM t, ErlNifBinary* out) {
  (void)env;
  cell* c = (cell*)t;
  if (c == NULL || c->kind != K_BIN) return 0;
  out->size = c->size;
  out->data = c->data;
  out->ref_bin = NULL;
  return 1;
}
int enif_is_identical(ERL_NIF_TERM a, ERL_NIF_TERM b) { return a == b; }

ErlNifResourceType* enif_open_resource_type(ErlNifEnv* env, const char* module, const char* name,
                                            ErlNifResourceDtor* dtor, ErlNifResourceFlags flags,
                                            ErlNifResourceFlags* tried) {
  (void)env;
  (void)module;
  (void)flags;
  ErlNifResourceType* t = (ErlNifResourceType*)calloc(1, sizeof(*t));
  snprintf(t->name, sizeof(t->name), "%s", name);
  t->dtor = dtor;
  if (tried) *tried = ERL_NIF_RT_CREATE;
  return t;
}
void* enif_alloc_resource(ErlNifResourceType* type, size_t size) {
  res_hdr* h = (res_hdr*)calloc(1, sizeof(res_hdr) + size);
  h->type = type;
  h->refs = 1;
  h->size = size;
  pthread_mutex_lock(&g_lock);
  ++g_live_resources;
  pthread_mutex_unlock(&g_lock);
  return h + 1;
}
void enif_release_resource(void* obj) { res_unref(obj); }
int enif_keep_resource(void* obj) {
  res_hdr* h = (res_hdr*)obj - 1;
  pthread_mutex_lock(&g_lock);
  ++h->refs;
  pthread_mutex_unlock(&g_lock);
  return 1;
}
ERL_NIF_TERM enif_make_resource(ErlNifEnv* env, void* obj) {
  res_hdr* h = (res_hdr*)obj - 1;
  pthread_mutex_lock(&g_lock);
  ++h->refs;                          /* the term keeps the object alive while its environment lives */
  pthread_mutex_unlock(&g_lock);
  cell* c = new_cell(env, K_RES);
  c->res = obj;
  return (ERL_NIF_TERM)c;
}
int enif_get_resource(ErlNifEnv* env, ERL_NIF_TERM t, ErlNifResourceType* type, void** out) {
  (void)env;
  cell* c = (cell*)t;
  if (c == NULL || c->kind != K_RES) return 0;
  res_hdr* h = (res_hdr*)c->res - 1;
  if (h->type != type) return 0;
  *out = c->res;
  return 1;
}

ErlNifMutex* enif_mutex_create(char* name) {
  (void)name;
  ErlNifMutex* m = (ErlNifMutex*)calloc(1, sizeof(*m));
  pthread_mutex_init(&m->m, NULL);
  return m;
}
void enif_mutex_destroy(ErlNifMutex* m) {
  pthread_mutex_destroy(&m->m);
  free(m);
}
void enif_mutex_lock(ErlNifMutex* m) { pthread_mutex_lock(&m->m); }
void enif_mutex_unlock(ErlNifMutex* m) { pthread_mutex_unlock(&m->m); }

/* ---- the VM side (used by the test through ctypes) -------------------------------------------------------- */
ErlNifEntry* nif_init(void);          /* defined by ERL_NIF_INIT in the NIF library linked with this file */

ErlNifEnv* mock_env_new(void) { return (ErlNifEnv*)calloc(1, sizeof(ErlNifEnv)); }

void mock_env_free(ErlNifEnv* env) {
  cell* c = env->cells;
  while (c) {
    cell* n = c->next;
    if (c->kind == K_RES) res_unref(c->res);
    free(c->elems);
    free(c->data);
    free(c);
    c = n;
  }
  free(env);
}

ERL_NIF_TERM mock_make_binary(ErlNifEnv* env, const void* data, size_t size) {
  ERL_NIF_TERM t;
  unsigned char* p = enif_make_new_binary(env, size, &t);
  if (size) memcpy(p, data, size);
  return t;
}
/* A term of another environment re-homed in `env` (what sending a term to another process does); resources gain a
 * reference, binaries are copied. */
ERL_NIF_TERM mock_copy_term(ErlNifEnv* env, ERL_NIF_TERM t) {
  cell* c = (cell*)t;
  switch (c->kind) {
    case K_ATOM: return t;
    case K_INT: return enif_make_int(env, (int)c->ival);
    case K_U64: return enif_make_uint64(env, c->uval);
    case K_BIN: return mock_make_binary(env, c->data, c->size);
    case K_RES: return enif_make_resource(env, c->res);
    case K_TUPLE: {
      ERL_NIF_TERM e[8];
      for (int i = 0; i < c->arity && i < 8; ++i) e[i] = mock_copy_term(env, (ERL_NIF_TERM)c->elems[i]);
      return make_tuple(env, c->arity, e);
    }
    default: return enif_make_badarg(env);
  }
}
int mock_term_kind(ERL_NIF_TERM t) { return ((cell*)t)->kind; }
int mock_tuple_arity(ERL_NIF_TERM t) { return ((cell*)t)->kind == K_TUPLE ? ((cell*)t)->arity : -1; }
ERL_NIF_TERM mock_tuple_elem(ERL_NIF_TERM t, int i) { return (ERL_NIF_TERM)((cell*)t)->elems[i]; }
const char* mock_atom_name(ERL_NIF_TERM t) { return ((cell*)t)->kind == K_ATOM ? ((cell*)t)->atom : NULL; }
const unsigned char* mock_binary_data(ERL_NIF_TERM t) { return ((cell*)t)->data; }
size_t mock_binary_size(ERL_NIF_TERM t) { return ((cell*)t)->size; }
unsigned long long mock_get_uint64(ERL_NIF_TERM t) {
  cell* c = (cell*)t;
  return c->kind == K_U64 ? c->uval : (unsigned long long)c->ival;
}
long mock_live_resources(void) { return g_live_resources; }

static void* g_priv = NULL;
/* on_load, as the VM does when the module is loaded; 0 on success */
int mock_load(void) {
  ErlNifEntry* e = nif_init();
  ErlNifEnv* env = mock_env_new();
  int rc = e->load ? e->load(env, &g_priv, enif_make_atom(env, "nil")) : 0;
  mock_env_free(env);
  return rc;
}
const char* mock_module_name(void) { return nif_init()->name; }
int mock_num_funcs(void) { return nif_init()->num_of_funcs; }
const char* mock_func_name(int i) { return nif_init()->funcs[i].name; }
unsigned mock_func_arity(int i) { return nif_init()->funcs[i].arity; }
unsigned mock_func_flags(int i) { return nif_init()->funcs[i].flags; }

/* Calls Module.name/argc; returns 0 (a NULL term) when the module exports no such function -- what `:erlang.nif_error`
 * would surface as an undefined function in a real VM. */
ERL_NIF_TERM mock_call(ErlNifEnv* env, const char* name, int argc, const ERL_NIF_TERM* argv) {
  ErlNifEntry* e = nif_init();
  for (int i = 0; i < e->num_of_funcs; ++i)
    if (strcmp(e->funcs[i].name, name) == 0 && (int)e->funcs[i].arity == argc) return e->funcs[i].fptr(env, argc, argv);
  return 0;
}
