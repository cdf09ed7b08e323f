/* minimpi.c -- see mpi.h. One process per rank; a full mesh of AF_UNIX stream socket pairs
 * inherited from the launcher (QB200_MINIMPI_RANK / _SIZE / _FDS). Messages are framed as
 * {source, tag, bytes} + payload; unmatched messages are queued (MPI's unexpected queue). */
#include "mpi.h"

#include <errno.h>
#include <poll.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <unistd.h>

#define MAX_RANKS 64
#define BCAST_TAG (-77001)
#define BARRIER_TAG (-77002)

typedef struct Msg {
  int source, tag, bytes;
  char *data;
  struct Msg *next;
} Msg;

static int g_rank = 0, g_size = 1, g_init = 0;
static int g_fd[MAX_RANKS];
static Msg *g_head = NULL, *g_tail = NULL;

static void die(const char *what) {
  fprintf(stderr, "minimpi[%d]: %s: %s\n", g_rank, what, strerror(errno));
  exit(-1);
}

static size_t type_size(MPI_Datatype t) {
  switch (t) {
    case MPI_BYTE:
    case MPI_CHAR: return 1;
    case MPI_INT:
    case MPI_UNSIGNED: return 4;
    case MPI_LONG_DOUBLE: return sizeof(long double);
    case MPI_DOUBLE:
    case MPI_UNSIGNED_LONG: return 8;
    default:
      fprintf(stderr, "minimpi: unknown datatype %d\n", t);
      exit(-1);
  }
}

static void write_all(int fd, const void *p, size_t n) {
  const char *c = (const char *)p;
  while (n) {
    const ssize_t k = write(fd, c, n);
    if (k < 0) {
      if (errno == EINTR) continue;
      die("write");
    }
    c += k;
    n -= (size_t)k;
  }
}

static int read_all(int fd, void *p, size_t n) {
  char *c = (char *)p;
  while (n) {
    const ssize_t k = read(fd, c, n);
    if (k < 0) {
      if (errno == EINTR) continue;
      die("read");
    }
    if (k == 0) return -1; /* peer closed */
    c += k;
    n -= (size_t)k;
  }
  return 0;
}

int MPI_Init_thread(int *argc, char ***argv, int required, int *provided) {
  (void)argc;
  (void)argv;
  const char *r = getenv("QB200_MINIMPI_RANK"), *s = getenv("QB200_MINIMPI_SIZE"),
             *f = getenv("QB200_MINIMPI_FDS");
  for (int i = 0; i < MAX_RANKS; i++) g_fd[i] = -1;
  if (r && s && f) {
    g_rank = atoi(r);
    g_size = atoi(s);
    if (g_size < 1 || g_size > MAX_RANKS) {
      fprintf(stderr, "minimpi: bad world size\n");
      exit(-1);
    }
    char *copy = strdup(f), *save = NULL;
    int i = 0;
    for (char *tok = strtok_r(copy, ",", &save); tok && i < g_size; tok = strtok_r(NULL, ",", &save))
      g_fd[i++] = atoi(tok);
    free(copy);
  }
  g_init = 1;
  if (provided) *provided = required < MPI_THREAD_FUNNELED ? required : MPI_THREAD_FUNNELED;
  return MPI_SUCCESS;
}

int MPI_Init(int *argc, char ***argv) { return MPI_Init_thread(argc, argv, MPI_THREAD_SINGLE, NULL); }

int MPI_Finalize(void) {
  for (int i = 0; i < g_size; i++)
    if (g_fd[i] >= 0) close(g_fd[i]);
  g_init = 0;
  return MPI_SUCCESS;
}

int MPI_Abort(MPI_Comm comm, int code) {
  (void)comm;
  exit(code ? code : -1);
}

int MPI_Comm_rank(MPI_Comm comm, int *rank) {
  (void)comm;
  *rank = g_rank;
  return MPI_SUCCESS;
}
int MPI_Comm_size(MPI_Comm comm, int *size) {
  (void)comm;
  *size = g_size;
  return MPI_SUCCESS;
}

static int send_bytes(const void *buf, size_t bytes, int dest, int tag) {
  if (dest < 0 || dest >= g_size || dest == g_rank || g_fd[dest] < 0) {
    fprintf(stderr, "minimpi[%d]: bad destination %d\n", g_rank, dest);
    exit(-1);
  }
  int hdr[3] = {g_rank, tag, (int)bytes};
  if (bytes && bytes <= 4096) { /* small messages: header and payload in one write */
    char tmp[sizeof(hdr) + 4096];
    memcpy(tmp, hdr, sizeof(hdr));
    memcpy(tmp + sizeof(hdr), buf, bytes);
    write_all(g_fd[dest], tmp, sizeof(hdr) + bytes);
    return MPI_SUCCESS;
  }
  write_all(g_fd[dest], hdr, sizeof(hdr));
  if (bytes) write_all(g_fd[dest], buf, bytes);
  return MPI_SUCCESS;
}

int MPI_Send(const void *buf, int count, MPI_Datatype type, int dest, int tag, MPI_Comm comm) {
  (void)comm;
  return send_bytes(buf, (size_t)count * type_size(type), dest, tag);
}

static Msg *read_msg(int fd) {
  int hdr[3];
  if (read_all(fd, hdr, sizeof(hdr)) != 0) return NULL;
  Msg *m = (Msg *)malloc(sizeof(Msg));
  m->source = hdr[0];
  m->tag = hdr[1];
  m->bytes = hdr[2];
  m->data = (char *)malloc(m->bytes > 0 ? (size_t)m->bytes : 1);
  m->next = NULL;
  if (m->bytes > 0 && read_all(fd, m->data, (size_t)m->bytes) != 0) {
    fprintf(stderr, "minimpi[%d]: peer %d closed mid-message\n", g_rank, m->source);
    exit(-1);
  }
  return m;
}

static int matches(const Msg *m, int source, int tag) {
  return (source == MPI_ANY_SOURCE || m->source == source) && (tag == MPI_ANY_TAG || m->tag == tag);
}

static int recv_bytes(void *buf, size_t cap, int source, int tag, MPI_Status *status) {
  for (;;) {
    Msg *prev = NULL;
    for (Msg *m = g_head; m; prev = m, m = m->next) {
      if (!matches(m, source, tag)) continue;
      if ((size_t)m->bytes > cap) {
        fprintf(stderr, "minimpi[%d]: message of %d bytes truncated (buffer %zu)\n", g_rank, m->bytes,
                cap);
        exit(-1);
      }
      memcpy(buf, m->data, (size_t)m->bytes);
      if (status) {
        status->MPI_SOURCE = m->source;
        status->MPI_TAG = m->tag;
        status->MPI_ERROR = MPI_SUCCESS;
        status->count_bytes = m->bytes;
      }
      if (prev) prev->next = m->next; else g_head = m->next;
      if (g_tail == m) g_tail = prev;
      free(m->data);
      free(m);
      return MPI_SUCCESS;
    }
    /* nothing queued matches. Receiving from ONE peer: read its next header and, if it is the
     * message wanted, the payload straight into the caller's buffer (the 256 KiB slice matrices
     * of *_slice_init_recv): no intermediate allocation, one copy less. */
    if (source != MPI_ANY_SOURCE && source >= 0 && source < g_size && g_fd[source] >= 0) {
      int hdr[3];
      if (read_all(g_fd[source], hdr, sizeof(hdr)) != 0) {
        fprintf(stderr, "minimpi[%d]: rank %d exited while a message was expected\n", g_rank, source);
        exit(-1);
      }
      if (tag == MPI_ANY_TAG || hdr[1] == tag) {
        if ((size_t)hdr[2] > cap) {
          fprintf(stderr, "minimpi[%d]: message of %d bytes truncated (buffer %zu)\n", g_rank, hdr[2], cap);
          exit(-1);
        }
        if (hdr[2] > 0 && read_all(g_fd[source], buf, (size_t)hdr[2]) != 0) {
          fprintf(stderr, "minimpi[%d]: peer %d closed mid-message\n", g_rank, source);
          exit(-1);
        }
        if (status) {
          status->MPI_SOURCE = hdr[0];
          status->MPI_TAG = hdr[1];
          status->MPI_ERROR = MPI_SUCCESS;
          status->count_bytes = hdr[2];
        }
        return MPI_SUCCESS;
      }
      /* some other tag: queue it and look again */
      Msg *m = (Msg *)malloc(sizeof(Msg));
      m->source = hdr[0];
      m->tag = hdr[1];
      m->bytes = hdr[2];
      m->data = (char *)malloc(m->bytes > 0 ? (size_t)m->bytes : 1);
      m->next = NULL;
      if (m->bytes > 0 && read_all(g_fd[source], m->data, (size_t)m->bytes) != 0) {
        fprintf(stderr, "minimpi[%d]: peer %d closed mid-message\n", g_rank, source);
        exit(-1);
      }
      if (g_tail) g_tail->next = m; else g_head = m;
      g_tail = m;
      continue;
    }
    /* any source: block for the next message from a candidate peer */
    struct pollfd pf[MAX_RANKS];
    int idx[MAX_RANKS], n = 0;
    for (int i = 0; i < g_size; i++) {
      if (g_fd[i] < 0) continue;
      if (source != MPI_ANY_SOURCE && i != source) continue;
      pf[n].fd = g_fd[i];
      pf[n].events = POLLIN;
      pf[n].revents = 0;
      idx[n++] = i;
    }
    if (n == 0) {
      fprintf(stderr, "minimpi[%d]: receive from nobody\n", g_rank);
      exit(-1);
    }
    if (poll(pf, (nfds_t)n, -1) < 0) {
      if (errno == EINTR) continue;
      die("poll");
    }
    for (int k = 0; k < n; k++) {
      if (!(pf[k].revents & (POLLIN | POLLHUP))) continue;
      Msg *m = read_msg(pf[k].fd);
      if (!m) {
        if (source != MPI_ANY_SOURCE) {
          fprintf(stderr, "minimpi[%d]: rank %d exited while a message was expected\n", g_rank, idx[k]);
          exit(-1);
        }
        close(g_fd[idx[k]]);
        g_fd[idx[k]] = -1;
        continue;
      }
      if (g_tail) g_tail->next = m; else g_head = m;
      g_tail = m;
    }
  }
}

int MPI_Recv(void *buf, int count, MPI_Datatype type, int source, int tag, MPI_Comm comm,
             MPI_Status *status) {
  (void)comm;
  return recv_bytes(buf, (size_t)count * type_size(type), source, tag, status);
}

int MPI_Bcast(void *buf, int count, MPI_Datatype type, int root, MPI_Comm comm) {
  (void)comm;
  const size_t bytes = (size_t)count * type_size(type);
  if (g_size == 1) return MPI_SUCCESS;
  if (g_rank == root) {
    for (int i = 0; i < g_size; i++)
      if (i != root) send_bytes(buf, bytes, i, BCAST_TAG);
    return MPI_SUCCESS;
  }
  return recv_bytes(buf, bytes, root, BCAST_TAG, NULL);
}

int MPI_Barrier(MPI_Comm comm) {
  (void)comm;
  char c = 0;
  if (g_size == 1) return MPI_SUCCESS;
  if (g_rank == 0) {
    for (int i = 1; i < g_size; i++) recv_bytes(&c, 1, i, BARRIER_TAG, NULL);
    for (int i = 1; i < g_size; i++) send_bytes(&c, 1, i, BARRIER_TAG);
  } else {
    send_bytes(&c, 1, 0, BARRIER_TAG);
    recv_bytes(&c, 1, 0, BARRIER_TAG, NULL);
  }
  return MPI_SUCCESS;
}
