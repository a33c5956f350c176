/*
  oracle/shim/mpi.h -- TEST INFRASTRUCTURE ONLY (never linked into the product).

  A minimal stand-in for <mpi.h> so that the UNMODIFIED reference sources under
  /root/reference/src compile in an image that has no MPI.  It provides exactly
  the 15 MPI entry points the reference calls.

  Ranks are forked processes (PCU_SHIM_NP environment variable, default 1) that
  share one anonymous mmap region; every collective is "write my slot, barrier,
  reduce all slots in rank order, barrier", so results are identical on every
  rank and independent of timing.  Implementation: mpi_shim.cpp.
*/
#ifndef PCU_ORACLE_MPI_SHIM_H
#define PCU_ORACLE_MPI_SHIM_H

#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef int MPI_Comm;
typedef int MPI_Datatype;
typedef int MPI_Op;
typedef int MPI_Info;
typedef long long MPI_Offset;
typedef struct { int unused; } MPI_Status;
typedef struct pcu_shim_file *MPI_File;

#define MPI_COMM_WORLD 0
#define MPI_COMM_SELF 1
#define MPI_COMM_NULL (-1)
#define MPI_SUCCESS 0

#define MPI_INT 1
#define MPI_DOUBLE 2
#define MPI_DOUBLE_COMPLEX 3
#define MPI_CHAR 4

#define MPI_SUM 1
#define MPI_MAX 2
#define MPI_MIN 3
#define MPI_BOR 4

#define MPI_IN_PLACE ((void *)1)
#define MPI_INFO_NULL 0
#define MPI_STATUS_IGNORE ((MPI_Status *)0)

#define MPI_MODE_RDONLY 1
#define MPI_MODE_WRONLY 2
#define MPI_MODE_CREATE 4

int MPI_Init(int *argc, char ***argv);
int MPI_Finalize(void);
int MPI_Comm_rank(MPI_Comm comm, int *rank);
int MPI_Comm_size(MPI_Comm comm, int *size);
double MPI_Wtime(void);
int MPI_Barrier(MPI_Comm comm);
int MPI_Allreduce(const void *sendbuf, void *recvbuf, int count,
                  MPI_Datatype type, MPI_Op op, MPI_Comm comm);
int MPI_Reduce(const void *sendbuf, void *recvbuf, int count,
               MPI_Datatype type, MPI_Op op, int root, MPI_Comm comm);
int MPI_Bcast(void *buf, int count, MPI_Datatype type, int root, MPI_Comm comm);
int MPI_Allgather(const void *sendbuf, int sendcount, MPI_Datatype sendtype,
                  void *recvbuf, int recvcount, MPI_Datatype recvtype,
                  MPI_Comm comm);

int MPI_File_open(MPI_Comm comm, const char *filename, int amode, MPI_Info info,
                  MPI_File *fh);
int MPI_File_close(MPI_File *fh);
int MPI_File_write(MPI_File fh, const void *buf, int count, MPI_Datatype type,
                   MPI_Status *status);
int MPI_File_read(MPI_File fh, void *buf, int count, MPI_Datatype type,
                  MPI_Status *status);
int MPI_File_set_view(MPI_File fh, MPI_Offset disp, MPI_Datatype etype,
                      MPI_Datatype filetype, const char *datarep,
                      MPI_Info info);
int MPI_File_write_at_all(MPI_File fh, MPI_Offset offset, const void *buf,
                          int count, MPI_Datatype type, MPI_Status *status);
int MPI_File_read_at_all(MPI_File fh, MPI_Offset offset, void *buf, int count,
                         MPI_Datatype type, MPI_Status *status);

#ifdef __cplusplus
}
#endif

#endif
