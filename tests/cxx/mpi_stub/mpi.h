// Minimal stand-in for <mpi.h>: lets the CPU test-suite compile-check the MPI overloads of
// include/superbblas.h (copy / contraction with an MPI_Comm) on machines without MPI.  Never used to run anything.
#pragma once
typedef int MPI_Comm;
typedef int MPI_Datatype;
#define MPI_COMM_WORLD 0
#define MPI_BYTE 1
#define MPI_INT 2
#define MPI_SUCCESS 0
inline int MPI_Init(int *, char ***) { return 0; }
inline int MPI_Finalize() { return 0; }
inline int MPI_Comm_rank(MPI_Comm, int *r) { *r = 0; return 0; }
inline int MPI_Comm_size(MPI_Comm, int *s) { *s = 1; return 0; }
inline int MPI_Bcast(void *, int, MPI_Datatype, int, MPI_Comm) { return 0; }
inline int MPI_Barrier(MPI_Comm) { return 0; }
