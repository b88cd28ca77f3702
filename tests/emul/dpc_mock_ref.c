/* tests/emul/dpc_mock_ref.c -- TEST DOUBLE, NOT PRODUCT CODE.
 *
 * A stand-in for libdynprog_cuda.so that answers the ticket API of include/dynprog_cuda.h (dpc_add / dpc_flush /
 * dpc_wait / dpc_result / dpc_pairs) with the compiled reference (oracle/_ref/libdynprog_ref.so, loaded with
 * RTLD_DEEPBIND so that its own Dynprog_* are used).  It exists so that the HOST-side scheduling of
 * gmap-gsnap_b200/host/dynprog_dropin.c (fibers, double-buffered batch contexts, ticket bookkeeping) can be
 * exercised by `pytest -m "not gpu"` inside a real gmap binary on a machine without a GPU:
 *     LD_LIBRARY_PATH=tests/emul/_mock oracle/_ref/gmap_cuda ...
 * Nothing in the product links or loads this file; on the GPU box the real library is used.
 */
#define _GNU_SOURCE
#include <dlfcn.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <libgen.h>
#include "../../include/dynprog_cuda.h"

typedef int (*ref_init_fn)(int, int, int, int, int, int);
typedef int (*ref_setup_fn)(const dpc_setup_t *);
typedef int (*ref_solve_fn)(const dpc_problem_t *, int, dpc_result_t *, dpc_pair_t *, int64_t, int64_t *);
static ref_init_fn ref_init_;
static ref_setup_fn ref_setup_;
static ref_solve_fn ref_solve_;

static int load_ref(void) {
  if (ref_solve_) return 0;
  Dl_info info;
  char path[4096];
  const char *env = getenv("DPC_MOCK_REF");
  if (env) snprintf(path, sizeof path, "%s", env);
  else {
    if (!dladdr((void *)load_ref, &info)) return -1;
    char tmp[4096];
    snprintf(tmp, sizeof tmp, "%s", info.dli_fname);
    snprintf(path, sizeof path, "%s/../../../oracle/_ref/libdynprog_ref.so", dirname(tmp));
  }
  void *h = dlopen(path, RTLD_NOW | RTLD_LOCAL | RTLD_DEEPBIND);
  if (!h) { fprintf(stderr, "dpc_mock_ref: %s\n", dlerror()); return -1; }
  ref_init_ = (ref_init_fn)dlsym(h, "ref_init");
  ref_setup_ = (ref_setup_fn)dlsym(h, "ref_setup");
  ref_solve_ = (ref_solve_fn)dlsym(h, "ref_solve");
  return ref_init_ && ref_setup_ && ref_solve_ ? 0 : -1;
}

struct dpc_ctx {
  dpc_problem_t *p; char **own; int n, cap;
  dpc_result_t *r; dpc_pair_t *pairs; int64_t *off; int64_t paircap;
  int flushed;
};

int dpc_init(int a, int b, int c, int d, int e, int mode) { return load_ref() ? DPC_ERR_CUDA : ref_init_(a, b, c, d, e, mode); }
int dpc_setup(const dpc_setup_t *s) { return load_ref() ? DPC_ERR_CUDA : ref_setup_(s); }
void dpc_term(void) {}
const char *dpc_strerror(int code) { (void)code; return "mock backend error"; }
int dpc_device_count(void) { return 1; }
int dpc_warmup(int device) { (void)device; return DPC_OK; }

dpc_ctx_t *dpc_ctx_new(int device) { (void)device; return (dpc_ctx_t *)calloc(1, sizeof(dpc_ctx_t)); }
void dpc_ctx_free(dpc_ctx_t *c) { if (c) { dpc_reset(c); free(c->p); free(c->own); free(c->r); free(c->pairs); free(c->off); free(c); } }

int dpc_reset(dpc_ctx_t *c) {
  for (int i = 0; i < 2 * c->n; i++) { free(c->own[i]); c->own[i] = NULL; }
  c->n = 0; c->flushed = 0;
  return DPC_OK;
}

/* forward pointers address the first character, "rev" pointers the last (include/dynprog_cuda.h) */
static const char *copy_seq(const char *s, int len, int rev, char **own) {
  if (!s || len <= 0) return s;
  char *buf = (char *)malloc((size_t)len + 2);
  memcpy(buf, rev ? s - (len - 1) : s, (size_t)len);
  buf[len] = 0;
  *own = buf;
  return rev ? buf + len - 1 : buf;
}

int dpc_add(dpc_ctx_t *c, const dpc_problem_t *q) {
  if (c->flushed) return DPC_ERR_STATE;
  if (c->n == c->cap) {
    int old = c->cap;
    c->cap = c->cap ? 2 * c->cap : 64;
    c->p = (dpc_problem_t *)realloc(c->p, c->cap * sizeof(dpc_problem_t));
    c->own = (char **)realloc(c->own, 2 * c->cap * sizeof(char *));
    memset(c->own + 2 * old, 0, 2 * (c->cap - old) * sizeof(char *));
  }
  dpc_problem_t *d = &c->p[c->n];
  *d = *q;
  const int five = q->kind == DPC_END5_GAP || q->kind == DPC_END5_SPLICEJUNCTION;
  d->seq1 = copy_seq(q->seq1, q->length1, five, &c->own[2 * c->n]);
  if (q->kind == DPC_CDNA_GAP) d->seq1R = copy_seq(q->seq1R, q->length1R, 1, &c->own[2 * c->n + 1]);
  if (q->kind == DPC_END5_SPLICEJUNCTION || q->kind == DPC_END3_SPLICEJUNCTION)
    d->seq1R = copy_seq(q->seq1R, q->length2, five, &c->own[2 * c->n + 1]);
  return c->n++;
}

int dpc_flush(dpc_ctx_t *c) {
  if (c->flushed) return DPC_ERR_STATE;
  c->flushed = 1;
  if (c->n == 0) return DPC_OK;
  c->r = (dpc_result_t *)realloc(c->r, c->n * sizeof(dpc_result_t));
  c->off = (int64_t *)realloc(c->off, (c->n + 1) * sizeof(int64_t));
  int64_t need = 0;
  for (int i = 0; i < c->n; i++) need += 2 * (int64_t)(c->p[i].length1 + c->p[i].length1R + c->p[i].length2 + c->p[i].length2R) + 64;
  if (need > c->paircap) { c->paircap = need; c->pairs = (dpc_pair_t *)realloc(c->pairs, need * sizeof(dpc_pair_t)); }
  return ref_solve_(c->p, c->n, c->r, c->pairs, c->paircap, c->off);
}

int dpc_wait(dpc_ctx_t *c) { return c->flushed ? DPC_OK : DPC_ERR_STATE; }

int dpc_result(dpc_ctx_t *c, int ticket, dpc_result_t *out) {
  if (!c->flushed || ticket < 0 || ticket >= c->n) return DPC_ERR_STATE;
  *out = c->r[ticket];
  return DPC_OK;
}

int dpc_pairs(dpc_ctx_t *c, int ticket, dpc_pair_t *out, int cap) {
  if (!c->flushed || ticket < 0 || ticket >= c->n) return DPC_ERR_STATE;
  int n = (int)(c->off[ticket + 1] - c->off[ticket]);
  if (n > cap) return DPC_ERR_ARG;
  memcpy(out, c->pairs + c->off[ticket], n * sizeof(dpc_pair_t));
  return n;
}
