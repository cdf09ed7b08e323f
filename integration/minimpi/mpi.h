/* mpi.h -- minimal single-node MPI for running the reference's generators where no MPI
 * implementation is installed (this image, and the GPU boxes).
 *
 * INTEGRATION-TEST INFRASTRUCTURE, not part of the product library. It implements exactly the
 * subset of MPI the reference's generation path uses (SURVEY.md section 2 "collective" census):
 * MPI_Init[_thread], MPI_Finalize, MPI_Comm_rank/size, blocking MPI_Send / MPI_Recv with tag
 * matching, MPI_ANY_SOURCE, MPI_ANY_TAG and MPI_Status, and MPI_Bcast, over Unix socket pairs
 * set up by the `minimpirun -np N` launcher. With a real MPI (mpicc/mpirun) none of this is
 * needed: the reference's own Makefile applies.
 */
#ifndef QB200_MINIMPI_H
#define QB200_MINIMPI_H

#ifdef __cplusplus
extern "C" {
#endif

typedef int MPI_Comm;
typedef int MPI_Datatype;
typedef struct {
  int MPI_SOURCE;
  int MPI_TAG;
  int MPI_ERROR;
  int count_bytes;
} MPI_Status;

#define MPI_SUCCESS 0
#define MPI_COMM_WORLD 0
#define MPI_ANY_SOURCE (-1)
#define MPI_ANY_TAG (-1)
#define MPI_STATUS_IGNORE ((MPI_Status *)0)

#define MPI_BYTE 1
#define MPI_CHAR 2
#define MPI_INT 3
#define MPI_UNSIGNED 4
#define MPI_LONG_DOUBLE 5
#define MPI_DOUBLE 6
#define MPI_UNSIGNED_LONG 7

#define MPI_THREAD_SINGLE 0
#define MPI_THREAD_FUNNELED 1
#define MPI_THREAD_SERIALIZED 2
#define MPI_THREAD_MULTIPLE 3

int MPI_Init(int *argc, char ***argv);
int MPI_Init_thread(int *argc, char ***argv, int required, int *provided);
int MPI_Finalize(void);
int MPI_Abort(MPI_Comm comm, int code);
int MPI_Comm_rank(MPI_Comm comm, int *rank);
int MPI_Comm_size(MPI_Comm comm, int *size);
int MPI_Barrier(MPI_Comm comm);
int MPI_Send(const void *buf, int count, MPI_Datatype type, int dest, int tag, MPI_Comm comm);
int MPI_Recv(void *buf, int count, MPI_Datatype type, int source, int tag, MPI_Comm comm,
             MPI_Status *status);
int MPI_Bcast(void *buf, int count, MPI_Datatype type, int root, MPI_Comm comm);

#ifdef __cplusplus
}
#endif

#endif /* QB200_MINIMPI_H */
