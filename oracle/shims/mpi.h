/* Link-only stand-in for <mpi.h>.
 *
 * TEST INFRASTRUCTURE ONLY.  The reference's parameters*.cpp and gmp_mpi.c
 * reference MPI_Bcast/MPI_Send/MPI_Recv for the broadcast of parameters; the
 * oracle never calls those functions (it is a single process), so they are
 * declared here and defined in oracle/ref_capi.cpp as functions that abort.
 */
#ifndef QUNUNDRUM_B200_SHIM_MPI_H
#define QUNUNDRUM_B200_SHIM_MPI_H

#ifdef __cplusplus
extern "C" {
#endif

typedef int MPI_Comm;
typedef int MPI_Datatype;
typedef struct {
  int MPI_SOURCE;
  int MPI_TAG;
  int MPI_ERROR;
} MPI_Status;

#define MPI_SUCCESS 0
#define MPI_COMM_WORLD 0
#define MPI_ANY_SOURCE (-1)
#define MPI_ANY_TAG (-1)
#define MPI_BYTE 1
#define MPI_CHAR 2
#define MPI_INT 3
#define MPI_UNSIGNED 4
#define MPI_LONG_DOUBLE 5
#define MPI_DOUBLE 6
#define MPI_UNSIGNED_LONG 7
#define MPI_STATUS_IGNORE ((MPI_Status *)0)

int MPI_Bcast(void *, int, MPI_Datatype, int, MPI_Comm);
int MPI_Send(const void *, int, MPI_Datatype, int, int, MPI_Comm);
int MPI_Recv(void *, int, MPI_Datatype, int, int, MPI_Comm, MPI_Status *);

#ifdef __cplusplus
}
#endif

#endif /* QUNUNDRUM_B200_SHIM_MPI_H */
