// h2d_ceiling.cu -- aggregate pinned host->device copy ceiling of the box: N GPUs copy at the same
// time, one host thread per GPU, for a few allocation / stream variants.  Evidence for DESIGN.md §5
// (what bounds the end-to-end sketch rate when ranks are added).
//   nvcc -O2 -o scripts/h2d_ceiling scripts/h2d_ceiling.cu -lpthread ; scripts/h2d_ceiling [bytes_per_gpu]
#include <cuda_runtime.h>
#include <pthread.h>
#include <sched.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>
#include <vector>

static double now() {
    timespec t;
    clock_gettime(CLOCK_MONOTONIC, &t);
    return t.tv_sec + 1e-9 * t.tv_nsec;
}

struct Job {
    int dev, nstreams, wc, reps;
    size_t bytes, chunk;
    pthread_barrier_t *bar;
    double gbs;
    char cpus[256];
};

static void bind_to_gpu_cpus(int dev, char *desc) {  // /sys/bus/pci/devices/<id>/local_cpulist
    char id[32], path[128], line[256] = "";
    cudaDeviceGetPCIBusId(id, sizeof id, dev);
    for (char *p = id; *p; p++) if (*p >= 'A' && *p <= 'Z') *p += 32;
    snprintf(path, sizeof path, "/sys/bus/pci/devices/%s/local_cpulist", id);
    FILE *f = fopen(path, "r");
    if (f) { if (!fgets(line, sizeof line, f)) line[0] = 0; fclose(f); }
    line[strcspn(line, "\n")] = 0;
    snprintf(desc, 256, "%s", line);
    cpu_set_t set;
    CPU_ZERO(&set);
    int any = 0;
    for (char *tok = strtok(line, ","); tok; tok = strtok(NULL, ",")) {
        int a, b;
        if (sscanf(tok, "%d-%d", &a, &b) == 2) { for (int c = a; c <= b; c++) CPU_SET(c, &set); any = 1; }
        else if (sscanf(tok, "%d", &a) == 1) { CPU_SET(a, &set); any = 1; }
    }
    if (any) sched_setaffinity(0, sizeof set, &set);
}

static void *run(void *arg) {
    Job *j = (Job *)arg;
    cudaSetDevice(j->dev);
    bind_to_gpu_cpus(j->dev, j->cpus);
    void *h, *d;
    cudaHostAlloc(&h, j->bytes, j->wc ? cudaHostAllocWriteCombined : cudaHostAllocDefault);
    memset(h, 1, j->bytes);  // first touch under the binding
    cudaMalloc(&d, j->bytes);
    std::vector<cudaStream_t> st(j->nstreams);
    for (auto &s : st) cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking);
    auto pass = [&]() {
        int k = 0;
        for (size_t o = 0; o < j->bytes; o += j->chunk, k++) {
            size_t n = j->bytes - o < j->chunk ? j->bytes - o : j->chunk;
            cudaMemcpyAsync((char *)d + o, (char *)h + o, n, cudaMemcpyHostToDevice, st[k % j->nstreams]);
        }
        for (auto &s : st) cudaStreamSynchronize(s);
    };
    pass();
    pthread_barrier_wait(j->bar);
    double t0 = now();
    for (int r = 0; r < j->reps; r++) pass();
    double t1 = now();
    pthread_barrier_wait(j->bar);
    j->gbs = 1e-9 * j->bytes * j->reps / (t1 - t0);
    cudaFree(d);
    cudaFreeHost(h);
    return NULL;
}

int main(int argc, char **argv) {
    size_t bytes = argc > 1 ? strtoull(argv[1], 0, 10) : (size_t)1 << 30;
    int ndev = 0;
    cudaGetDeviceCount(&ndev);
    printf("{\"gpus_visible\": %d, \"bytes_per_gpu\": %zu, \"runs\": [\n", ndev, bytes);
    int first = 1;
    for (int n = 1; n <= ndev; n *= 2)
        for (int wc = 0; wc < 2; wc++)
            for (int ns = 1; ns <= 2; ns++)
                for (size_t chunk : {(size_t)16 << 20, (size_t)256 << 20}) {
                    pthread_barrier_t bar;
                    pthread_barrier_init(&bar, NULL, n);
                    std::vector<Job> jobs(n);
                    std::vector<pthread_t> th(n);
                    for (int i = 0; i < n; i++) {
                        jobs[i] = Job{i, ns, wc, 6, bytes, chunk, &bar, 0.0, ""};
                        pthread_create(&th[i], NULL, run, &jobs[i]);
                    }
                    double sum = 0, mn = 1e30;
                    for (int i = 0; i < n; i++) {
                        pthread_join(th[i], NULL);
                        sum += jobs[i].gbs;
                        if (jobs[i].gbs < mn) mn = jobs[i].gbs;
                    }
                    printf("%s{\"gpus\": %d, \"write_combined\": %d, \"streams\": %d, \"chunk_mib\": %zu, "
                           "\"aggregate_gbs\": %.1f, \"min_per_gpu_gbs\": %.1f, \"gpu0_local_cpus\": \"%s\"}",
                           first ? "" : ",\n", n, wc, ns, chunk >> 20, sum, mn, jobs[0].cpus);
                    first = 0;
                    fflush(stdout);
                    pthread_barrier_destroy(&bar);
                }
    printf("\n]}\n");
    return 0;
}
