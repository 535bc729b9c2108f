// sample_sim.cpp — runs csrc/sample.cu's kernel on the CPU emulator (cusim.h).  TEST INFRASTRUCTURE.
#include "../../ntransformer_b200/csrc/kernels_internal.h"      // before cusim.h (launch_k names cudaLaunchConfig_t::gridDim)
#include <random>
#include <vector>
#include "cusim.h"
#include "../../ntransformer_b200/csrc/sample.cu"

using namespace nt::b200;

extern "C" {

// Same arguments as nt_sample_token / the reference shim's ref_sample_token: the uniform variate is the first draw of
// std::mt19937(seed) through std::uniform_real_distribution<float>, as in Sampler::sample (sampler.cpp:105-106).
// Returns the token id, -1 when the settings are not covered by the GPU kernel, -2 on an emulator failure.
int sample_sim(const float* logits, int n, float temperature, int top_k, float top_p, float repeat_penalty, int repeat_window,
               const int* recent, int n_recent, unsigned long long seed) {
    if (!sample_topk_supported(n, temperature, top_k)) return -1;
    std::vector<float> l(logits, logits + n);
    const int window = repeat_penalty > 1.0f ? std::min(n_recent, repeat_window) : 0;
    std::vector<int> win(recent + (n_recent - window), recent + n_recent);
    std::mt19937 rng;
    rng.seed(seed);
    std::uniform_real_distribution<float> dist(0.0f, 1.0f);
    const float r = dist(rng);
    int out = -3;
    SampleParams p{l.data(), n, temperature, top_p, repeat_penalty, r, top_k, win.data(), window, &out};
    cusim::g_failed = false;
    if (!cusim::launch(1, ST, 0, 0, [p]() { sample_topk_kernel(p); }) || cusim::g_failed.load()) return -2;
    return out;
}

}  // extern "C"
