// h2d_ceiling.cu — what the platform's pinned-memory copies can do with NO kernel in between: one process per GPU, all at
// once (fork), the chunked three-stream copy loop of the library's host-buffer calls (fqtk_b200_copy_ceiling).
// bench.py reports the same measurement in its line as e2e.ceiling; this is the stand-alone form.
// Build: nvcc -O3 -o tools/h2d_ceiling tools/h2d_ceiling.cu -Ifqtk_b200/../include -Lfqtk_b200 -lfqtk_b200 -Xlinker -rpath -Xlinker $PWD/fqtk_b200
// usage: tools/h2d_ceiling [n_gpus] [reads_per_gpu] [bytes_in_per_read] [bytes_out_per_read]
#include <cstdio>
#include <cstdlib>
#include <sys/wait.h>
#include <unistd.h>

#include "../include/fqtk_b200.h"

int main(int argc, char** argv) {
    const int n_gpus = argc > 1 ? atoi(argv[1]) : 1;
    const unsigned long long reads = argc > 2 ? strtoull(argv[2], nullptr, 10) : (128ull << 20);
    const unsigned long long bin = argc > 3 ? strtoull(argv[3], nullptr, 10) : 16, bout = argc > 4 ? strtoull(argv[4], nullptr, 10) : 4;
    int pipes[64][2];
    for (int g = 0; g < n_gpus; g++) {
        if (pipe(pipes[g]) != 0) return 1;
        if (fork() == 0) {  // one process per GPU, like the bench's ranks (CUDA is first touched after the fork)
            double sec = 0;
            const int rc = fqtk_b200_copy_ceiling(g, reads * bin, reads * bout, 32ull << 20, 5, &sec);
            if (rc != FQTK_B200_OK) {
                fprintf(stderr, "gpu %d: %s\n", g, fqtk_b200_last_error());
                sec = -1;
            }
            if (write(pipes[g][1], &sec, sizeof sec) != (ssize_t)sizeof sec) _exit(2);
            _exit(0);
        }
    }
    double worst = 0;
    for (int g = 0; g < n_gpus; g++) {
        double sec = -1;
        if (read(pipes[g][0], &sec, sizeof sec) != (ssize_t)sizeof sec || sec <= 0) return 1;
        printf("gpu %d: %.2f GB/s in + %.2f GB/s out = %.1f Mreads/s\n", g, reads * bin / sec / 1e9, reads * bout / sec / 1e9, reads / sec / 1e6);
        if (sec > worst) worst = sec;
    }
    while (wait(nullptr) > 0) {}
    printf("%d GPU(s), %llu reads each of %llu B in + %llu B out: %.1f Mreads/s in total (slowest GPU)\n", n_gpus, reads, bin, bout,
           n_gpus * reads / worst / 1e6);
    return 0;
}
