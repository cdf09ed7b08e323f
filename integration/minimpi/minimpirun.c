/* minimpirun -np N program [args...] -- launcher of the minimal MPI (see mpi.h). */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/socket.h>
#include <sys/wait.h>
#include <unistd.h>

#define MAX_RANKS 64

int main(int argc, char **argv) {
  if (argc < 4 || strcmp(argv[1], "-np") != 0) {
    fprintf(stderr, "usage: minimpirun -np N program [args...]\n");
    return 2;
  }
  const int n = atoi(argv[2]);
  if (n < 1 || n > MAX_RANKS) {
    fprintf(stderr, "minimpirun: N must be in [1, %d]\n", MAX_RANKS);
    return 2;
  }
  static int fd[MAX_RANKS][MAX_RANKS];
  for (int i = 0; i < n; i++)
    for (int j = 0; j < n; j++) fd[i][j] = -1;
  for (int i = 0; i < n; i++)
    for (int j = i + 1; j < n; j++) {
      int sv[2];
      if (socketpair(AF_UNIX, SOCK_STREAM, 0, sv) != 0) {
        perror("socketpair");
        return 1;
      }
      /* room for several slice messages in flight (default 208 KiB < one 256 KiB slice) */
      for (int e = 0; e < 2; e++) {
        int sz = 8 << 20;
        setsockopt(sv[e], SOL_SOCKET, SO_SNDBUF, &sz, sizeof(sz));
        setsockopt(sv[e], SOL_SOCKET, SO_RCVBUF, &sz, sizeof(sz));
      }
      fd[i][j] = sv[0];
      fd[j][i] = sv[1];
    }
  pid_t pids[MAX_RANKS];
  for (int r = 0; r < n; r++) {
    const pid_t pid = fork();
    if (pid < 0) {
      perror("fork");
      return 1;
    }
    if (pid == 0) {
      char buf[32], fds[MAX_RANKS * 12] = "";
      for (int i = 0; i < n; i++)
        for (int j = 0; j < n; j++)
          if (i != r && fd[i][j] >= 0) close(fd[i][j]);
      for (int j = 0; j < n; j++) {
        snprintf(buf, sizeof(buf), "%s%d", j ? "," : "", fd[r][j]);
        strcat(fds, buf);
      }
      snprintf(buf, sizeof(buf), "%d", r);
      setenv("QB200_MINIMPI_RANK", buf, 1);
      snprintf(buf, sizeof(buf), "%d", n);
      setenv("QB200_MINIMPI_SIZE", buf, 1);
      setenv("QB200_MINIMPI_FDS", fds, 1);
      execvp(argv[3], argv + 3);
      perror("execvp");
      _exit(127);
    }
    pids[r] = pid;
  }
  for (int i = 0; i < n; i++)
    for (int j = 0; j < n; j++)
      if (fd[i][j] >= 0) close(fd[i][j]);
  int rc = 0;
  for (int r = 0; r < n; r++) {
    int st = 0;
    waitpid(pids[r], &st, 0);
    if (!WIFEXITED(st) || WEXITSTATUS(st) != 0) rc = 1;
  }
  return rc;
}
