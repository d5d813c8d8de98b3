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
/* point-to-point: not reachable from the single-rank oracle driver */
static inline int nts_shim_mpi_unreachable_(void) { abort(); return 1; }
static inline int MPI_Send(const void *, int, MPI_Datatype, int, int, MPI_Comm) { return nts_shim_mpi_unreachable_(); }
static inline int MPI_Issend(const void *, int, MPI_Datatype, int, int, MPI_Comm, MPI_Request *) { return nts_shim_mpi_unreachable_(); }
static inline int MPI_Recv(void *, int, MPI_Datatype, int, int, MPI_Comm, MPI_Status *) { return nts_shim_mpi_unreachable_(); }
static inline int MPI_Probe(int, int, MPI_Comm, MPI_Status *) { return nts_shim_mpi_unreachable_(); }
static inline int MPI_Iprobe(int, int, MPI_Comm, int *, MPI_Status *) { return nts_shim_mpi_unreachable_(); }
static inline int MPI_Improbe(int, int, MPI_Comm, int *, MPI_Message *, MPI_Status *) { return nts_shim_mpi_unreachable_(); }
static inline int MPI_Mrecv(void *, int, MPI_Datatype, MPI_Message *, MPI_Status *) { return nts_shim_mpi_unreachable_(); }
static inline int MPI_Get_count(const MPI_Status *, MPI_Datatype, int *) { return nts_shim_mpi_unreachable_(); }
static inline int MPI_Test(MPI_Request *, int *, MPI_Status *) { return nts_shim_mpi_unreachable_(); }
static inline int MPI_Wait(MPI_Request *, MPI_Status *) { return nts_shim_mpi_unreachable_(); }
#endif
