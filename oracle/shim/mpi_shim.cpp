/*
  oracle/shim/mpi_shim.cpp -- TEST INFRASTRUCTURE ONLY (never linked into the
  product).  Implementation of the mpi.h stand-in used to build the unmodified
  reference as a CPU oracle / CPU baseline.

  Process model: MPI_Init forks PCU_SHIM_NP-1 children (default: single rank).
  All ranks share one MAP_SHARED|MAP_ANONYMOUS region holding a sense-reversing
  barrier and one payload slot per rank.  Collectives:
      write own slot -> barrier -> every rank combines slots 0..P-1 in rank
      order -> barrier
  so every rank obtains a bit-identical result (the reference relies on that
  only loosely, via root-compute + Bcast).
*/
#include "mpi.h"

#include <errno.h>
#include <fcntl.h>
#include <sched.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/mman.h>
#include <sys/time.h>
#include <sys/wait.h>
#include <time.h>
#include <unistd.h>

#include <atomic>

namespace {

const size_t SLOT_BYTES = 4u << 20;  // payload per rank (>= 100x100 doubles)

struct Shared {
  std::atomic<int> count;
  std::atomic<int> sense;
  char pad[56];
};

Shared *g_sh = nullptr;
char *g_slots = nullptr;
int g_rank = 0, g_size = 1;
int g_local_sense = 0;
pid_t g_children[1024];

size_t type_size(MPI_Datatype t) {
  switch (t) {
    case MPI_INT:
      return sizeof(int);
    case MPI_DOUBLE:
      return sizeof(double);
    case MPI_DOUBLE_COMPLEX:
      return 2 * sizeof(double);
    case MPI_CHAR:
      return 1;
  }
  fprintf(stderr, "mpi shim: unknown datatype %d\n", t);
  abort();
}

void barrier() {
  if (g_size == 1) return;
  g_local_sense = !g_local_sense;
  if (g_sh->count.fetch_add(1) == g_size - 1) {
    g_sh->count.store(0);
    g_sh->sense.store(g_local_sense);
  } else {
    int spins = 0;
    while (g_sh->sense.load() != g_local_sense) {
      if (++spins > 2000) {
        sched_yield();
      }
    }
  }
}

char *slot(int r) { return g_slots + (size_t)r * SLOT_BYTES; }

template <class T>
void combine(T *acc, const T *in, int count, MPI_Op op) {
  for (int i = 0; i < count; i++) {
    switch (op) {
      case MPI_SUM:
        acc[i] = acc[i] + in[i];
        break;
      case MPI_MAX:
        if (in[i] > acc[i]) acc[i] = in[i];
        break;
      case MPI_MIN:
        if (in[i] < acc[i]) acc[i] = in[i];
        break;
      default:
        break;
    }
  }
}

void combine_any(void *acc, const void *in, int count, MPI_Datatype t,
                 MPI_Op op) {
  if (t == MPI_DOUBLE) {
    combine((double *)acc, (const double *)in, count, op);
  } else if (t == MPI_DOUBLE_COMPLEX) {
    combine((double *)acc, (const double *)in, 2 * count, op);  // SUM only
  } else if (t == MPI_INT) {
    if (op == MPI_BOR) {
      int *a = (int *)acc;
      const int *b = (const int *)in;
      for (int i = 0; i < count; i++) a[i] |= b[i];
    } else {
      combine((int *)acc, (const int *)in, count, op);
    }
  }
}

}  // namespace

struct pcu_shim_file {
  int fd;
  MPI_Offset disp;
  size_t etype_size;
};

extern "C" {

int MPI_Init(int *, char ***) {
  const char *np = getenv("PCU_SHIM_NP");
  g_size = np ? atoi(np) : 1;
  if (g_size < 1) g_size = 1;
  if (g_size > 1024) g_size = 1024;
  size_t bytes = sizeof(Shared) + (size_t)g_size * SLOT_BYTES;
  void *mem = mmap(nullptr, bytes, PROT_READ | PROT_WRITE,
                   MAP_SHARED | MAP_ANONYMOUS, -1, 0);
  if (mem == MAP_FAILED) {
    perror("mpi shim: mmap");
    abort();
  }
  g_sh = new (mem) Shared;
  g_sh->count.store(0);
  g_sh->sense.store(0);
  g_slots = (char *)mem + sizeof(Shared);
  g_rank = 0;
  fflush(stdout);
  fflush(stderr);
  for (int r = 1; r < g_size; r++) {
    pid_t pid = fork();
    if (pid == 0) {
      g_rank = r;
      break;
    }
    g_children[r] = pid;
  }
  return MPI_SUCCESS;
}

int MPI_Finalize(void) {
  barrier();
  fflush(stdout);
  fflush(stderr);
  if (g_rank != 0) {
    _exit(0);
  }
  for (int r = 1; r < g_size; r++) {
    int status = 0;
    waitpid(g_children[r], &status, 0);
  }
  return MPI_SUCCESS;
}

int MPI_Comm_rank(MPI_Comm comm, int *rank) {
  *rank = (comm == MPI_COMM_SELF) ? 0 : g_rank;
  return MPI_SUCCESS;
}

int MPI_Comm_size(MPI_Comm comm, int *size) {
  *size = (comm == MPI_COMM_SELF) ? 1 : g_size;
  return MPI_SUCCESS;
}

double MPI_Wtime(void) {
  struct timespec ts;
  clock_gettime(CLOCK_MONOTONIC, &ts);
  return (double)ts.tv_sec + 1e-9 * (double)ts.tv_nsec;
}

int MPI_Barrier(MPI_Comm comm) {
  if (comm != MPI_COMM_SELF) barrier();
  return MPI_SUCCESS;
}

int MPI_Allreduce(const void *sendbuf, void *recvbuf, int count,
                  MPI_Datatype type, MPI_Op op, MPI_Comm comm) {
  size_t bytes = (size_t)count * type_size(type);
  const void *src = (sendbuf == MPI_IN_PLACE) ? recvbuf : sendbuf;
  if (g_size == 1 || comm == MPI_COMM_SELF) {
    if (src != recvbuf) memmove(recvbuf, src, bytes);
    return MPI_SUCCESS;
  }
  if (bytes > SLOT_BYTES) {
    fprintf(stderr, "mpi shim: payload too large (%zu bytes)\n", bytes);
    abort();
  }
  memcpy(slot(g_rank), src, bytes);
  barrier();
  memcpy(recvbuf, slot(0), bytes);
  for (int r = 1; r < g_size; r++) {
    combine_any(recvbuf, slot(r), count, type, op);
  }
  barrier();
  return MPI_SUCCESS;
}

int MPI_Reduce(const void *sendbuf, void *recvbuf, int count, MPI_Datatype type,
               MPI_Op op, int root, MPI_Comm comm) {
  if (g_size == 1 || comm == MPI_COMM_SELF) {
    return MPI_Allreduce(sendbuf, recvbuf, count, type, op, comm);
  }
  // Non-root ranks may pass recvbuf == NULL; reduce into scratch for them.
  size_t bytes = (size_t)count * type_size(type);
  void *tmp = recvbuf;
  bool own = false;
  if (g_rank != root || recvbuf == nullptr) {
    tmp = malloc(bytes ? bytes : 1);
    own = true;
  }
  const void *src = (sendbuf == MPI_IN_PLACE) ? recvbuf : sendbuf;
  MPI_Allreduce(src, tmp, count, type, op, comm);
  if (own) free(tmp);
  return MPI_SUCCESS;
}

int MPI_Bcast(void *buf, int count, MPI_Datatype type, int root,
              MPI_Comm comm) {
  if (g_size == 1 || comm == MPI_COMM_SELF) return MPI_SUCCESS;
  size_t bytes = (size_t)count * type_size(type);
  if (bytes > SLOT_BYTES) {
    fprintf(stderr, "mpi shim: bcast payload too large\n");
    abort();
  }
  if (g_rank == root) memcpy(slot(root), buf, bytes);
  barrier();
  if (g_rank != root) memcpy(buf, slot(root), bytes);
  barrier();
  return MPI_SUCCESS;
}

int MPI_Allgather(const void *sendbuf, int sendcount, MPI_Datatype sendtype,
                  void *recvbuf, int, MPI_Datatype, MPI_Comm comm) {
  size_t bytes = (size_t)sendcount * type_size(sendtype);
  if (g_size == 1 || comm == MPI_COMM_SELF) {
    memmove(recvbuf, sendbuf, bytes);
    return MPI_SUCCESS;
  }
  memcpy(slot(g_rank), sendbuf, bytes);
  barrier();
  for (int r = 0; r < g_size; r++) {
    memcpy((char *)recvbuf + (size_t)r * bytes, slot(r), bytes);
  }
  barrier();
  return MPI_SUCCESS;
}

/* ---- MPI-IO over POSIX files (checkpoint path; off the hot path) ---- */

int MPI_File_open(MPI_Comm, const char *filename, int amode, MPI_Info,
                  MPI_File *fh) {
  int flags = 0;
  if (amode & MPI_MODE_WRONLY) flags |= O_WRONLY;
  if (amode & MPI_MODE_CREATE) flags |= O_CREAT;
  if (amode & MPI_MODE_RDONLY) flags |= O_RDONLY;
  int fd = open(filename, flags, 0644);
  if (fd < 0) {
    *fh = nullptr;
    return 1;
  }
  pcu_shim_file *f = (pcu_shim_file *)malloc(sizeof(pcu_shim_file));
  f->fd = fd;
  f->disp = 0;
  f->etype_size = 1;
  *fh = f;
  return MPI_SUCCESS;
}

int MPI_File_close(MPI_File *fh) {
  if (fh && *fh) {
    close((*fh)->fd);
    free(*fh);
    *fh = nullptr;
  }
  barrier();
  return MPI_SUCCESS;
}

int MPI_File_write(MPI_File fh, const void *buf, int count, MPI_Datatype type,
                   MPI_Status *) {
  ssize_t n = write(fh->fd, buf, (size_t)count * type_size(type));
  return n < 0 ? 1 : MPI_SUCCESS;
}

int MPI_File_read(MPI_File fh, void *buf, int count, MPI_Datatype type,
                  MPI_Status *) {
  ssize_t n = read(fh->fd, buf, (size_t)count * type_size(type));
  return n < 0 ? 1 : MPI_SUCCESS;
}

int MPI_File_set_view(MPI_File fh, MPI_Offset disp, MPI_Datatype etype,
                      MPI_Datatype, const char *, MPI_Info) {
  fh->disp = disp;
  fh->etype_size = type_size(etype);
  return MPI_SUCCESS;
}

int MPI_File_write_at_all(MPI_File fh, MPI_Offset offset, const void *buf,
                          int count, MPI_Datatype type, MPI_Status *) {
  off_t pos = (off_t)(fh->disp + offset * (MPI_Offset)fh->etype_size);
  ssize_t n = pwrite(fh->fd, buf, (size_t)count * type_size(type), pos);
  return n < 0 ? 1 : MPI_SUCCESS;
}

int MPI_File_read_at_all(MPI_File fh, MPI_Offset offset, void *buf, int count,
                         MPI_Datatype type, MPI_Status *) {
  off_t pos = (off_t)(fh->disp + offset * (MPI_Offset)fh->etype_size);
  ssize_t n = pread(fh->fd, buf, (size_t)count * type_size(type), pos);
  return n < 0 ? 1 : MPI_SUCCESS;
}

/* ---- METIS stand-in: identity ordering ---- */

int METIS_SetDefaultOptions(int *options) {
  for (int i = 0; i < 40; i++) options[i] = -1;
  return 1;
}

int METIS_NodeND(int *nvtxs, int *, int *, int *, int *, int *perm,
                 int *iperm) {
  for (int i = 0; i < *nvtxs; i++) {
    perm[i] = i;
    iperm[i] = i;
  }
  return 1;
}

}  // extern "C"
