/* Single-node MPI subset over POSIX shared memory -- ONLY so that the UNMODIFIED reference sources
 * (/root/reference/{dmnist,dcifar10}/.../*.cpp) can be compiled and timed on a box that has no MPI.
 * Implements exactly the calls the reference makes (SURVEY.md section 2.5 inventory):
 *   Init / Finalize / Comm_rank / Comm_size / Wtime / Alloc_mem / Win_create / Win_lock / Win_unlock /
 *   Win_flush / Put / Allreduce(IN_PLACE, SUM) / Issend / Recv / Wait   (+ Barrier, Free_mem, Win_free)
 * Ranks are OS processes started by baseline/shim/mpirun (env EGMPI_JOB / EGMPI_RANK / EGMPI_SIZE).
 * Not part of the product: the product's transport is eventgrad_b200/csrc (CUDA peer memory).
 */
#ifndef EGMPI_MPI_H
#define EGMPI_MPI_H
#include <stddef.h>
#ifdef __cplusplus
extern "C" {
#endif

typedef int MPI_Comm;
typedef int MPI_Datatype;
typedef int MPI_Op;
typedef int MPI_Info;
typedef int MPI_Win;
typedef int MPI_Request;
typedef long MPI_Aint;
typedef struct { int MPI_SOURCE, MPI_TAG, MPI_ERROR; } MPI_Status;

#define MPI_SUCCESS 0
#define MPI_COMM_WORLD 0
#define MPI_INFO_NULL 0
#define MPI_IN_PLACE ((void *)1)
#define MPI_STATUS_IGNORE ((MPI_Status *)0)
#define MPI_LOCK_EXCLUSIVE 1
#define MPI_LOCK_SHARED 2
#define MPI_SUM 1

/* datatype id = element size in the low byte, kind in the next */
#define MPI_CHAR 0x0101
#define MPI_UNSIGNED_CHAR 0x0201
#define MPI_SHORT 0x0302
#define MPI_INT 0x0404
#define MPI_LONG 0x0508
#define MPI_FLOAT 0x0604
#define MPI_DOUBLE 0x0708

int MPI_Init(int *argc, char ***argv);
int MPI_Finalize(void);
int MPI_Comm_rank(MPI_Comm comm, int *rank);
int MPI_Comm_size(MPI_Comm comm, int *size);
double MPI_Wtime(void);
int MPI_Barrier(MPI_Comm comm);
int MPI_Alloc_mem(MPI_Aint size, MPI_Info info, void *baseptr);
int MPI_Free_mem(void *base);
int MPI_Win_create(void *base, MPI_Aint size, int disp_unit, MPI_Info info, MPI_Comm comm, MPI_Win *win);
int MPI_Win_free(MPI_Win *win);
int MPI_Win_lock(int lock_type, int rank, int assert_, MPI_Win win);
int MPI_Win_unlock(int rank, MPI_Win win);
int MPI_Win_flush(int rank, MPI_Win win);
int MPI_Put(const void *origin, int origin_count, MPI_Datatype origin_type, int target_rank,
            MPI_Aint target_disp, int target_count, MPI_Datatype target_type, MPI_Win win);
int MPI_Allreduce(const void *sendbuf, void *recvbuf, int count, MPI_Datatype type, MPI_Op op, MPI_Comm comm);
int MPI_Issend(const void *buf, int count, MPI_Datatype type, int dest, int tag, MPI_Comm comm, MPI_Request *req);
int MPI_Recv(void *buf, int count, MPI_Datatype type, int source, int tag, MPI_Comm comm, MPI_Status *status);
int MPI_Wait(MPI_Request *req, MPI_Status *status);

#ifdef __cplusplus
}
#endif
#endif
