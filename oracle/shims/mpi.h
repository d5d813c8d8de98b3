/* oracle/shims/mpi.h -- single-rank, in-process stand-in for <mpi.h>.
 *
 * TEST INFRASTRUCTURE ONLY. Lets the reference's CPU path (core/, comm/) be
 * compiled from /root/reference without an MPI installation. Every sample
 * toolkit of the reference runs 1 rank, so collectives
 * are identities and point-to-point is never reached by the oracle driver. */
#ifndef NTS_ORACLE_SHIM_MPI_H
#define NTS_ORACLE_SHIM_MPI_H
#include <stddef.h>
#include <stdlib.h>
#include <string.h>
#include <sys/time.h>

typedef int MPI_Comm;
typedef int MPI_Datatype;
typedef int MPI_Op;
typedef int MPI_Request;
typedef struct MPI_Status { int MPI_SOURCE; int MPI_TAG; int MPI_ERROR; int count_; } MPI_Status;
typedef struct MPI_Message_ { int unused; } *MPI_Message;

#define MPI_COMM_WORLD 0
#define MPI_SUCCESS 0
#define MPI_ANY_SOURCE (-1)
#define MPI_ANY_TAG (-1)
#define MPI_STATUS_IGNORE ((MPI_Status *)0)
#define MPI_IN_PLACE ((void *)1)
#define MPI_THREAD_SINGLE 0
#define MPI_THREAD_FUNNELED 1
#define MPI_THREAD_SERIALIZED 2
#define MPI_THREAD_MULTIPLE 3

/* datatype handle = element size in bytes (enough for the identity collectives) */
#define MPI_CHAR 1
#define MPI_UNSIGNED_CHAR 1
#define MPI_UINT8_T 1
#define MPI_INT 4
#define MPI_UNSIGNED 4
#define MPI_FLOAT 4
#define MPI_LONG 8
#define MPI_UNSIGNED_LONG 8
#define MPI_DOUBLE 8
#define MPI_SUM 0
#define MPI_MIN 1
#define MPI_MAX 2

static inline int MPI_Init_thread(int *, char ***, int required, int *provided) { if (provided) *provided = required; return 0; }
static inline int MPI_Initialized(int *flag) { *flag = 1; return 0; }
static inline int MPI_Finalize(void) { return 0; }
static inline int MPI_Abort(MPI_Comm, int code) { exit(code); return 0; }
static inline int MPI_Comm_rank(MPI_Comm, int *r) { *r = 0; return 0; }
static inline int MPI_Comm_size(MPI_Comm, int *s) { *s = 1; return 0; }
static inline int MPI_Barrier(MPI_Comm) { return 0; }
static inline double MPI_Wtime(void) { struct timeval tv; gettimeofday(&tv, 0); return tv.tv_sec + tv.tv_usec * 1e-6; }
static inline int MPI_Allreduce(const void *s, void *r, int n, MPI_Datatype t, MPI_Op, MPI_Comm) {
  if (s != MPI_IN_PLACE && s != r) memcpy(r, s, (size_t)n * (size_t)t);
  return 0;
}
static inline int MPI_Bcast(void *, int, MPI_Datatype, int, MPI_Comm) { return 0; }
/* point-to-point: rank 0 sending to itself. Graph::load_directed ships every edge chunk to its owner with
 * MPI_Send / MPI_Probe / MPI_Get_count / MPI_Recv from a sender and a receiver thread (core/graph.hpp:1337-1417);
 * with one rank the owner is always self, so an in-process tagged mailbox is sufficient. */
#ifdef __cplusplus
#include <condition_variable>
#include <deque>
#include <mutex>
#include <vector>
struct nts_shim_msg { int tag; std::vector<char> data; };
struct nts_shim_mailbox { std::mutex m; std::condition_variable cv; std::deque<nts_shim_msg> q; };
inline nts_shim_mailbox &nts_shim_box() { static nts_shim_mailbox b; return b; }
inline int MPI_Send(const void *buf, int count, MPI_Datatype t, int, int tag, MPI_Comm) {
  nts_shim_mailbox &b = nts_shim_box();
  nts_shim_msg msg; msg.tag = tag; msg.data.assign((const char *)buf, (const char *)buf + (size_t)count * (size_t)t);
  { std::lock_guard<std::mutex> g(b.m); b.q.push_back(std::move(msg)); }
  b.cv.notify_all();
  return 0;
}
inline int MPI_Issend(const void *buf, int count, MPI_Datatype t, int d, int tag, MPI_Comm c, MPI_Request *r) { if (r) *r = 0; return MPI_Send(buf, count, t, d, tag, c); }
inline int nts_shim_find(nts_shim_mailbox &b, int tag) {
  for (size_t i = 0; i < b.q.size(); i++) if (tag == MPI_ANY_TAG || b.q[i].tag == tag) return (int)i;
  return -1;
}
inline int MPI_Probe(int, int tag, MPI_Comm, MPI_Status *st) {
  nts_shim_mailbox &b = nts_shim_box();
  std::unique_lock<std::mutex> g(b.m);
  int i;
  b.cv.wait(g, [&] { return (i = nts_shim_find(b, tag)) >= 0; });
  if (st) { st->MPI_SOURCE = 0; st->MPI_TAG = b.q[i].tag; st->MPI_ERROR = 0; st->count_ = (int)b.q[i].data.size(); }
  return 0;
}
inline int MPI_Iprobe(int, int tag, MPI_Comm, int *flag, MPI_Status *st) {
  nts_shim_mailbox &b = nts_shim_box();
  std::lock_guard<std::mutex> g(b.m);
  int i = nts_shim_find(b, tag);
  *flag = i >= 0;
  if (i >= 0 && st) { st->MPI_SOURCE = 0; st->MPI_TAG = b.q[i].tag; st->MPI_ERROR = 0; st->count_ = (int)b.q[i].data.size(); }
  return 0;
}
inline int MPI_Get_count(const MPI_Status *st, MPI_Datatype t, int *count) { *count = st->count_ / t; return 0; }
inline int MPI_Recv(void *buf, int count, MPI_Datatype t, int, int tag, MPI_Comm, MPI_Status *st) {
  nts_shim_mailbox &b = nts_shim_box();
  std::unique_lock<std::mutex> g(b.m);
  int i;
  b.cv.wait(g, [&] { return (i = nts_shim_find(b, tag)) >= 0; });
  size_t n = b.q[i].data.size(), cap = (size_t)count * (size_t)t;
  memcpy(buf, b.q[i].data.data(), n < cap ? n : cap);
  if (st) { st->MPI_SOURCE = 0; st->MPI_TAG = b.q[i].tag; st->MPI_ERROR = 0; st->count_ = (int)n; }
  b.q.erase(b.q.begin() + i);
  return 0;
}
inline int MPI_Improbe(int s, int tag, MPI_Comm c, int *flag, MPI_Message *, MPI_Status *st) { return MPI_Iprobe(s, tag, c, flag, st); }
inline int MPI_Mrecv(void *buf, int count, MPI_Datatype t, MPI_Message *, MPI_Status *st) { return MPI_Recv(buf, count, t, 0, MPI_ANY_TAG, 0, st); }
inline int MPI_Test(MPI_Request *, int *flag, MPI_Status *) { *flag = 1; return 0; }
inline int MPI_Wait(MPI_Request *, MPI_Status *) { return 0; }
#endif
#endif
