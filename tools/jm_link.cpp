/*
 * tools/jm_link.cpp -- the host link's ceiling for a whole box, barrier-synchronised: one thread per GPU, every phase
 * (H2D only / D2H only / both at once, optionally from write-combined host memory) starts on all GPUs together and is
 * device-timed per GPU (jmc_link_probe).  Plain C++ on include/jmc_cuda.h.
 *
 *   tools/jm_link [--gpus N] [--mb 64] [--ms 400 | --copies K]      one JSON line: per-GPU and whole-box GB/s per phase
 * Default: every GPU copies for the same 400 ms window (with a fixed number of copies the faster GPUs finish early and
 * the slower ones then run alone, which overstates the concurrent rate).
 */
#include <pthread.h>

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <thread>
#include <vector>

#include "jmc_cuda.h"

struct Phase { const char *name; int mode; };
static const Phase PHASES[] = { { "h2d_only", 1 }, { "d2h_only", 2 }, { "bidirectional", 3 }, { "h2d_only_wc", 5 }, { "bidirectional_wc", 7 } };
constexpr int NPH = sizeof(PHASES) / sizeof(PHASES[0]);

int main(int argc, char **argv)
{
    int gpus = jmc_device_count(), mb = 64, copies = -400;       /* copies < 0: a window of that many milliseconds */
    for (int i = 1; i < argc; i++) {
        std::string a = argv[i];
        if (a == "--gpus" && i + 1 < argc) gpus = atoi(argv[++i]);
        else if (a == "--mb" && i + 1 < argc) mb = atoi(argv[++i]);
        else if (a == "--copies" && i + 1 < argc) copies = atoi(argv[++i]);
        else if (a == "--ms" && i + 1 < argc) copies = -atoi(argv[++i]);
    }
    if (gpus < 1) { fprintf(stderr, "no CUDA device\n"); return 1; }
    pthread_barrier_t bar;
    pthread_barrier_init(&bar, nullptr, (unsigned)gpus);
    std::vector<std::vector<jmc_link_rates>> res((size_t)gpus, std::vector<jmc_link_rates>(NPH));
    std::vector<std::string> err((size_t)gpus);
    std::vector<std::thread> th;
    for (int g = 0; g < gpus; g++)
        th.emplace_back([&, g]() {
            jmc_ctx *c = nullptr;
            if (jmc_ctx_create(g, &c) != JMC_OK) { err[(size_t)g] = jmc_last_error(); }
            for (int p = 0; p < NPH; p++) {
                pthread_barrier_wait(&bar);                       /* every GPU enters the phase together */
                if (c && jmc_link_probe(c, (size_t)mb << 20, copies, PHASES[p].mode, &res[(size_t)g][(size_t)p]) != JMC_OK) err[(size_t)g] = jmc_last_error();
                pthread_barrier_wait(&bar);
            }
            if (c) jmc_ctx_destroy(c);
        });
    for (auto &t : th) t.join();
    printf("{\"n_gpus\": %d, \"mb_per_copy\": %d, \"copies\": %d", gpus, mb, copies);
    for (int p = 0; p < NPH; p++) {
        double up = 0, down = 0;
        printf(", \"%s\": {\"per_gpu_h2d_gbs\": [", PHASES[p].name);
        for (int g = 0; g < gpus; g++) { printf("%s%.2f", g ? ", " : "", res[(size_t)g][(size_t)p].h2d_gbs); up += res[(size_t)g][(size_t)p].h2d_gbs; }
        printf("], \"per_gpu_d2h_gbs\": [");
        for (int g = 0; g < gpus; g++) { printf("%s%.2f", g ? ", " : "", res[(size_t)g][(size_t)p].d2h_gbs); down += res[(size_t)g][(size_t)p].d2h_gbs; }
        printf("], \"box_h2d_gbs\": %.2f, \"box_d2h_gbs\": %.2f}", up, down);
    }
    for (int g = 0; g < gpus; g++) if (!err[(size_t)g].empty()) printf(", \"error_gpu%d\": \"%s\"", g, err[(size_t)g].c_str());
    printf("}\n");
    return 0;
}
