/* Single-process MPI shim -- see stub/mpi.h.  TEST INFRASTRUCTURE ONLY. */
#include "mpi.h"
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

static int g_inited = 0, g_finalized = 0;
static int g_next_comm = 3;
static size_t g_extent[4096];
static int g_next_type = MPI_SHIM_FIRST_DERIVED;
static int g_next_op = MPI_SHIM_FIRST_USER_OP;
/* groups: handle -> size (0 or 1); handle 1 == EMPTY */
static int g_group_size[4096] = {0, 0};
static int g_next_group = 2;

static size_t type_size(MPI_Datatype t) {
  switch (t) {
    case MPI_CHAR: case MPI_UNSIGNED_CHAR: case MPI_BYTE: return 1;
    case MPI_SHORT: return sizeof(short);
    case MPI_INT: case MPI_UNSIGNED: return sizeof(int);
    case MPI_LONG: case MPI_UNSIGNED_LONG: return sizeof(long);
    case MPI_LONG_LONG_INT: case MPI_UNSIGNED_LONG_LONG: return sizeof(long long);
    case MPI_FLOAT: return 4;
    case MPI_DOUBLE: return 8;
    case MPI_LONG_DOUBLE: return sizeof(long double);
    case MPI_COMPLEX: case MPI_C_FLOAT_COMPLEX: return 8;
    case MPI_DOUBLE_COMPLEX: case MPI_C_DOUBLE_COMPLEX: return 16;
    case MPI_FLOAT_INT: return sizeof(struct { float a; int b; });
    case MPI_DOUBLE_INT: return sizeof(struct { double a; int b; });
    case MPI_LONG_INT: return sizeof(struct { long a; int b; });
    case MPI_2INT: return 2 * sizeof(int);
    default:
      if (t >= MPI_SHIM_FIRST_DERIVED && t < 4096) return g_extent[t];
      fprintf(stderr, "mpi_shim: unknown datatype %d\n", t);
      abort();
  }
}
static void cpy(const void* s, void* r, size_t bytes) {
  if (s != MPI_IN_PLACE && s != r && bytes) memmove(r, s, bytes);
}

int MPI_Init(int* a, char*** b) { (void)a; (void)b; g_inited = 1; return 0; }
int MPI_Init_thread(int* a, char*** b, int req, int* prov) { (void)a; (void)b; g_inited = 1; if (prov) *prov = req; return 0; }
int MPI_Initialized(int* f) { *f = g_inited; return 0; }
int MPI_Finalize(void) { g_finalized = 1; return 0; }
int MPI_Finalized(int* f) { *f = g_finalized; return 0; }
int MPI_Query_thread(int* p) { *p = MPI_THREAD_MULTIPLE; return 0; }
int MPI_Abort(MPI_Comm c, int code) { (void)c; fprintf(stderr, "MPI_Abort(%d)\n", code); abort(); }
double MPI_Wtime(void) { struct timespec ts; clock_gettime(CLOCK_MONOTONIC, &ts); return ts.tv_sec + 1e-9 * ts.tv_nsec; }
int MPI_Error_string(int code, char* s, int* len) { *len = snprintf(s, MPI_MAX_ERROR_STRING, "mpi_shim error %d", code); return 0; }

int MPI_Comm_rank(MPI_Comm c, int* r) { (void)c; *r = 0; return 0; }
int MPI_Comm_size(MPI_Comm c, int* s) { (void)c; *s = 1; return 0; }
int MPI_Comm_dup(MPI_Comm c, MPI_Comm* n) { (void)c; *n = g_next_comm++; return 0; }
int MPI_Comm_split(MPI_Comm c, int color, int key, MPI_Comm* n) { (void)c; (void)key; *n = (color == MPI_UNDEFINED) ? MPI_COMM_NULL : g_next_comm++; return 0; }
int MPI_Comm_create(MPI_Comm c, MPI_Group g, MPI_Comm* n) { (void)c; *n = (g_group_size[g] > 0) ? g_next_comm++ : MPI_COMM_NULL; return 0; }
int MPI_Comm_free(MPI_Comm* c) { *c = MPI_COMM_NULL; return 0; }
static int new_group(int size) { int g = g_next_group++; if (g >= 4096) { g = 2; g_next_group = 3; } g_group_size[g] = size; return g; }
int MPI_Comm_group(MPI_Comm c, MPI_Group* g) { (void)c; *g = new_group(1); return 0; }
int MPI_Comm_compare(MPI_Comm a, MPI_Comm b, int* r) { *r = (a == b) ? MPI_IDENT : MPI_CONGRUENT; return 0; }
MPI_Comm MPI_Comm_f2c(int f) { return (MPI_Comm)f; }
int MPI_Comm_set_errhandler(MPI_Comm c, MPI_Errhandler e) { (void)c; (void)e; return 0; }
int MPI_Errhandler_set(MPI_Comm c, MPI_Errhandler e) { (void)c; (void)e; return 0; }
int MPI_Cart_create(MPI_Comm c, int nd, const int* d, const int* p, int re, MPI_Comm* n) { (void)c; (void)nd; (void)d; (void)p; (void)re; *n = g_next_comm++; return 0; }
int MPI_Cart_sub(MPI_Comm c, const int* rem, MPI_Comm* n) { (void)c; (void)rem; *n = g_next_comm++; return 0; }

int MPI_Group_rank(MPI_Group g, int* r) { *r = g_group_size[g] > 0 ? 0 : MPI_UNDEFINED; return 0; }
int MPI_Group_size(MPI_Group g, int* s) { *s = g_group_size[g]; return 0; }
int MPI_Group_incl(MPI_Group g, int n, const int* ranks, MPI_Group* o) { (void)g; (void)ranks; *o = n > 0 ? new_group(1) : MPI_GROUP_EMPTY; return 0; }
int MPI_Group_excl(MPI_Group g, int n, const int* ranks, MPI_Group* o) { (void)ranks; *o = (n > 0 || g_group_size[g] == 0) ? MPI_GROUP_EMPTY : new_group(1); return 0; }
int MPI_Group_union(MPI_Group a, MPI_Group b, MPI_Group* o) { *o = (g_group_size[a] || g_group_size[b]) ? new_group(1) : MPI_GROUP_EMPTY; return 0; }
int MPI_Group_difference(MPI_Group a, MPI_Group b, MPI_Group* o) { *o = (g_group_size[a] && !g_group_size[b]) ? new_group(1) : MPI_GROUP_EMPTY; return 0; }
int MPI_Group_free(MPI_Group* g) { *g = MPI_GROUP_NULL; return 0; }
int MPI_Group_compare(MPI_Group a, MPI_Group b, int* r) { *r = (g_group_size[a] == g_group_size[b]) ? MPI_IDENT : MPI_UNEQUAL; return 0; }
int MPI_Group_translate_ranks(MPI_Group a, int n, const int* ra, MPI_Group b, int* rb) { (void)a; for (int i = 0; i < n; ++i) rb[i] = g_group_size[b] > 0 ? ra[i] : MPI_UNDEFINED; return 0; }

int MPI_Barrier(MPI_Comm c) { (void)c; return 0; }
int MPI_Bcast(void* b, int n, MPI_Datatype t, int root, MPI_Comm c) { (void)b; (void)n; (void)t; (void)root; (void)c; return 0; }
int MPI_Ibcast(void* b, int n, MPI_Datatype t, int root, MPI_Comm c, MPI_Request* r) { (void)b; (void)n; (void)t; (void)root; (void)c; *r = 0; return 0; }
int MPI_Gather(const void* s, int sn, MPI_Datatype st, void* r, int rn, MPI_Datatype rt, int root, MPI_Comm c) { (void)rn; (void)rt; (void)root; (void)c; cpy(s, r, sn * type_size(st)); return 0; }
int MPI_Igather(const void* s, int sn, MPI_Datatype st, void* r, int rn, MPI_Datatype rt, int root, MPI_Comm c, MPI_Request* q) { *q = 0; return MPI_Gather(s, sn, st, r, rn, rt, root, c); }
int MPI_Gatherv(const void* s, int sn, MPI_Datatype st, void* r, const int* rns, const int* displs, MPI_Datatype rt, int root, MPI_Comm c) { (void)rns; (void)root; (void)c; if (s != MPI_IN_PLACE) memmove((char*)r + displs[0] * type_size(rt), s, sn * type_size(st)); return 0; }
int MPI_Scatter(const void* s, int sn, MPI_Datatype st, void* r, int rn, MPI_Datatype rt, int root, MPI_Comm c) { (void)sn; (void)st; (void)root; (void)c; if (r != MPI_IN_PLACE) cpy(s, r, rn * type_size(rt)); return 0; }
int MPI_Scatterv(const void* s, const int* sns, const int* displs, MPI_Datatype st, void* r, int rn, MPI_Datatype rt, int root, MPI_Comm c) { (void)sns; (void)root; (void)c; if (r != MPI_IN_PLACE) memmove(r, (const char*)s + displs[0] * type_size(st), rn * type_size(rt)); return 0; }
int MPI_Allgather(const void* s, int sn, MPI_Datatype st, void* r, int rn, MPI_Datatype rt, MPI_Comm c) { (void)rn; (void)rt; (void)c; cpy(s, r, sn * type_size(st)); return 0; }
int MPI_Allgatherv(const void* s, int sn, MPI_Datatype st, void* r, const int* rns, const int* displs, MPI_Datatype rt, MPI_Comm c) { (void)rns; (void)c; if (s != MPI_IN_PLACE) memmove((char*)r + displs[0] * type_size(rt), s, sn * type_size(st)); return 0; }
int MPI_Alltoall(const void* s, int sn, MPI_Datatype st, void* r, int rn, MPI_Datatype rt, MPI_Comm c) { (void)rn; (void)rt; (void)c; cpy(s, r, sn * type_size(st)); return 0; }
int MPI_Alltoallv(const void* s, const int* sns, const int* sd, MPI_Datatype st, void* r, const int* rns, const int* rd, MPI_Datatype rt, MPI_Comm c) { (void)rns; (void)c; if (s != MPI_IN_PLACE) memmove((char*)r + rd[0] * type_size(rt), (const char*)s + sd[0] * type_size(st), sns[0] * type_size(st)); return 0; }
int MPI_Reduce(const void* s, void* r, int n, MPI_Datatype t, MPI_Op op, int root, MPI_Comm c) { (void)op; (void)root; (void)c; cpy(s, r, n * type_size(t)); return 0; }
int MPI_Allreduce(const void* s, void* r, int n, MPI_Datatype t, MPI_Op op, MPI_Comm c) { (void)op; (void)c; cpy(s, r, n * type_size(t)); return 0; }
int MPI_Reduce_scatter(const void* s, void* r, const int* ns, MPI_Datatype t, MPI_Op op, MPI_Comm c) { (void)op; (void)c; cpy(s, r, ns[0] * type_size(t)); return 0; }
int MPI_Reduce_scatter_block(const void* s, void* r, int n, MPI_Datatype t, MPI_Op op, MPI_Comm c) { (void)op; (void)c; cpy(s, r, n * type_size(t)); return 0; }
int MPI_Scan(const void* s, void* r, int n, MPI_Datatype t, MPI_Op op, MPI_Comm c) { (void)op; (void)c; cpy(s, r, n * type_size(t)); return 0; }

/* ---- self point-to-point: FIFO of buffered sends + deferred receives ---- */
typedef struct Msg { int comm, tag; size_t bytes; void* data; struct Msg* next; } Msg;
static Msg *q_head = 0, *q_tail = 0;
typedef struct { int used, comm, tag, done; void* buf; size_t cap; size_t got; } PendingRecv;
static PendingRecv g_recv[1024];

static void enqueue(const void* b, size_t bytes, int tag, int comm) {
  Msg* m = (Msg*)malloc(sizeof(Msg));
  m->comm = comm; m->tag = tag; m->bytes = bytes; m->next = 0;
  m->data = malloc(bytes ? bytes : 1);
  memcpy(m->data, b, bytes);
  if (q_tail) q_tail->next = m; else q_head = m;
  q_tail = m;
}
static int dequeue(void* b, size_t cap, int tag, int comm, size_t* got, int* gtag) {
  Msg *p = 0, *m = q_head;
  while (m) {
    if (m->comm == comm && (tag == MPI_ANY_TAG || m->tag == tag)) {
      size_t n = m->bytes < cap ? m->bytes : cap;
      memcpy(b, m->data, n);
      if (got) *got = n;
      if (gtag) *gtag = m->tag;
      if (p) p->next = m->next; else q_head = m->next;
      if (q_tail == m) q_tail = p;
      free(m->data); free(m);
      return 1;
    }
    p = m; m = m->next;
  }
  return 0;
}
static void fill_status(MPI_Status* st, int tag, size_t bytes) {
  if (st) { st->MPI_SOURCE = 0; st->MPI_TAG = tag; st->MPI_ERROR = 0; st->count_bytes = (int)bytes; }
}
int MPI_Send(const void* b, int n, MPI_Datatype t, int dst, int tag, MPI_Comm c) { (void)dst; enqueue(b, n * type_size(t), tag, c); return 0; }
int MPI_Isend(const void* b, int n, MPI_Datatype t, int dst, int tag, MPI_Comm c, MPI_Request* r) { *r = 0; return MPI_Send(b, n, t, dst, tag, c); }
int MPI_Issend(const void* b, int n, MPI_Datatype t, int dst, int tag, MPI_Comm c, MPI_Request* r) { *r = 0; return MPI_Send(b, n, t, dst, tag, c); }
int MPI_Irsend(const void* b, int n, MPI_Datatype t, int dst, int tag, MPI_Comm c, MPI_Request* r) { *r = 0; return MPI_Send(b, n, t, dst, tag, c); }
int MPI_Recv(void* b, int n, MPI_Datatype t, int src, int tag, MPI_Comm c, MPI_Status* st) {
  (void)src; size_t got = 0; int gtag = tag;
  if (!dequeue(b, n * type_size(t), tag, c, &got, &gtag)) { fprintf(stderr, "mpi_shim: MPI_Recv would deadlock\n"); abort(); }
  fill_status(st, gtag, got); return 0;
}
int MPI_Irecv(void* b, int n, MPI_Datatype t, int src, int tag, MPI_Comm c, MPI_Request* r) {
  (void)src;
  for (int i = 1; i < 1024; ++i) if (!g_recv[i].used) {
    g_recv[i].used = 1; g_recv[i].comm = c; g_recv[i].tag = tag; g_recv[i].buf = b;
    g_recv[i].cap = n * type_size(t); g_recv[i].done = 0; g_recv[i].got = 0;
    *r = i; return 0;
  }
  fprintf(stderr, "mpi_shim: too many pending receives\n"); abort();
}
static int try_complete(int i, MPI_Status* st) {
  PendingRecv* p = &g_recv[i]; int gtag = p->tag;
  if (!p->done && dequeue(p->buf, p->cap, p->tag, p->comm, &p->got, &gtag)) p->done = 1;
  if (p->done) { fill_status(st, gtag, p->got); p->used = 0; return 1; }
  return 0;
}
int MPI_Wait(MPI_Request* r, MPI_Status* st) {
  if (*r > 0) { if (!try_complete(*r, st)) { fprintf(stderr, "mpi_shim: MPI_Wait would deadlock\n"); abort(); } *r = 0; }
  return 0;
}
int MPI_Waitall(int n, MPI_Request* r, MPI_Status* st) { for (int i = 0; i < n; ++i) MPI_Wait(&r[i], st ? &st[i] : 0); return 0; }
int MPI_Test(MPI_Request* r, int* flag, MPI_Status* st) { if (*r > 0) { *flag = try_complete(*r, st); if (*flag) *r = 0; } else *flag = 1; return 0; }
int MPI_Iprobe(int src, int tag, MPI_Comm c, int* flag, MPI_Status* st) {
  (void)src; *flag = 0;
  for (Msg* m = q_head; m; m = m->next) if (m->comm == c && (tag == MPI_ANY_TAG || m->tag == tag)) { *flag = 1; fill_status(st, m->tag, m->bytes); break; }
  return 0;
}
int MPI_Sendrecv(const void* s, int sn, MPI_Datatype st_, int dst, int stag, void* r, int rn, MPI_Datatype rt, int src, int rtag, MPI_Comm c, MPI_Status* st) {
  (void)dst; (void)src; (void)stag;
  size_t sb = sn * type_size(st_), rb = rn * type_size(rt);
  size_t n = sb < rb ? sb : rb;
  if (s != r && n) memmove(r, s, n);
  fill_status(st, rtag, n); return 0;
}
int MPI_Sendrecv_replace(void* b, int n, MPI_Datatype t, int dst, int stag, int src, int rtag, MPI_Comm c, MPI_Status* st) {
  (void)b; (void)dst; (void)stag; (void)src; (void)c; fill_status(st, rtag, n * type_size(t)); return 0;
}
int MPI_Get_count(const MPI_Status* st, MPI_Datatype t, int* n) { *n = (int)(st->count_bytes / type_size(t)); return 0; }
int MPI_Get_address(const void* p, MPI_Aint* a) { *a = (MPI_Aint)p; return 0; }

static int new_type(size_t extent) { int t = g_next_type++; if (t >= 4096) { fprintf(stderr, "mpi_shim: type table full\n"); abort(); } g_extent[t] = extent; return t; }
int MPI_Type_contiguous(int n, MPI_Datatype t, MPI_Datatype* o) { *o = new_type(n * type_size(t)); return 0; }
int MPI_Type_create_struct(int n, const int* bl, const MPI_Aint* disp, const MPI_Datatype* ts, MPI_Datatype* o) {
  size_t ext = 0; for (int i = 0; i < n; ++i) { size_t e = disp[i] + bl[i] * type_size(ts[i]); if (e > ext) ext = e; }
  *o = new_type(ext); return 0;
}
int MPI_Type_create_resized(MPI_Datatype t, MPI_Aint lb, MPI_Aint ext, MPI_Datatype* o) { (void)t; (void)lb; *o = new_type((size_t)ext); return 0; }
int MPI_Type_commit(MPI_Datatype* t) { (void)t; return 0; }
int MPI_Type_free(MPI_Datatype* t) { *t = MPI_DATATYPE_NULL; return 0; }
int MPI_Op_create(MPI_User_function* f, int commute, MPI_Op* op) { (void)f; (void)commute; *op = g_next_op++; return 0; }
int MPI_Op_free(MPI_Op* op) { *op = MPI_OP_NULL; return 0; }
