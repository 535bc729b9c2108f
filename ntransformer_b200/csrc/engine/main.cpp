// main.cpp — CLI with the reference's flag surface (src/main.cpp:6-174).  The streaming family of flags
// (--streaming, --draft-model, --draft-k, --self-spec, --early-exit, --skip-threshold, --requant-q4k,
// --delta-model) belongs to the PCIe/NVMe streaming subsystem that this resident engine drops: they are
// rejected with an explanation instead of being silently ignored.
#include "engine.h"
#include <cstring>
#include <string>

using namespace nt::b200;

static void usage(const char* prog) {
    fprintf(stderr, "NTransformer (B200 resident engine)\n\nUsage: %s [options] -m <model.gguf>\n\nOptions:\n", prog);
    fprintf(stderr, "  -m, --model <path>       Path to GGUF model file (required)\n");
    fprintf(stderr, "  -p, --prompt <text>      Prompt text (default: interactive mode)\n");
    fprintf(stderr, "  -n, --n-tokens <int>     Max tokens to generate (default: 256)\n");
    fprintf(stderr, "  -t, --temperature <float> Temperature (default: 0.7)\n");
    fprintf(stderr, "  --top-k <int>            Top-K sampling (default: 40)\n");
    fprintf(stderr, "  --top-p <float>          Top-P nucleus sampling (default: 0.9)\n");
    fprintf(stderr, "  --repeat-penalty <float> Repeat penalty (default: 1.1)\n");
    fprintf(stderr, "  -c, --ctx-size <int>     Context size (default: 4096)\n");
    fprintf(stderr, "  --seed <int>             Random seed (default: 42)\n");
    fprintf(stderr, "  --host-sampler           Sample (penalty, top-k, top-p) on the host like the reference (default: on the GPU)\n");
    fprintf(stderr, "  --megakernel             Decode each token with one persistent kernel [opt-in]\n");
    fprintf(stderr, "  --benchmark              Run benchmark mode\n");
    fprintf(stderr, "  --chat                   Interactive chat mode\n");
    fprintf(stderr, "  -v, --verbose            Verbose output\n");
    fprintf(stderr, "  -h, --help               Show this help\n");
    fprintf(stderr, "Not supported (resident engine; every model lives in HBM): --streaming --draft-model --draft-k\n"
                    "  --self-spec --early-exit --skip-threshold --requant-q4k --delta-model\n");
}

int main(int argc, char** argv) {
    std::string model_path, prompt;
    int max_context = 4096;
    bool bench = false, chat = false;
    GenerateConfig cfg;
    cfg.verbose = true;
    for (int i = 1; i < argc; i++) {
        std::string a = argv[i];
        auto val = [&]() -> const char* { return (i + 1 < argc) ? argv[++i] : nullptr; };
        if (a == "-h" || a == "--help") { usage(argv[0]); return 0; }
        else if (a == "-m" || a == "--model") { if (auto v = val()) model_path = v; }
        else if (a == "-p" || a == "--prompt") { if (auto v = val()) prompt = v; }
        else if (a == "-n" || a == "--n-tokens") { if (auto v = val()) cfg.max_tokens = std::stoi(v); }
        else if (a == "-t" || a == "--temperature") { if (auto v = val()) cfg.temperature = std::stof(v); }
        else if (a == "--top-k") { if (auto v = val()) cfg.top_k = std::stoi(v); }
        else if (a == "--top-p") { if (auto v = val()) cfg.top_p = std::stof(v); }
        else if (a == "--repeat-penalty") { if (auto v = val()) cfg.repeat_penalty = std::stof(v); }
        else if (a == "--seed") { if (auto v = val()) cfg.seed = std::stoull(v); }
        else if (a == "-c" || a == "--ctx-size") { if (auto v = val()) max_context = std::stoi(v); }
        else if (a == "--gpu-sampler") cfg.gpu_sampler = true;
        else if (a == "--host-sampler") cfg.gpu_sampler = false;
        else if (a == "--megakernel") setenv("NT_B200_MEGAKERNEL", "1", 1);
        else if (a == "--benchmark") bench = true;
        else if (a == "--chat") chat = true;
        else if (a == "-v" || a == "--verbose") cfg.verbose = true;
        else if (a == "--streaming" || a == "--self-spec" || a == "--requant-q4k" || a == "--draft-model" || a == "--draft-k" ||
                 a == "--early-exit" || a == "--skip-threshold" || a == "--delta-model") {
            fprintf(stderr, "Error: %s is not supported (resident engine: no layer streaming, all weights live in HBM)\n", a.c_str());
            return 1;
        } else { fprintf(stderr, "Unknown option: %s\n", a.c_str()); usage(argv[0]); return 1; }
    }
    if (model_path.empty()) { fprintf(stderr, "Error: model path required (-m)\n\n"); usage(argv[0]); return 1; }
    Engine engine;
    if (!engine.load(model_path, max_context)) { fprintf(stderr, "Failed to load model: %s\n", model_path.c_str()); return 1; }
    if (bench) engine.benchmark(prompt.empty() ? "The meaning of life is" : prompt, cfg.max_tokens);
    else if (chat || prompt.empty()) engine.chat(cfg);
    else engine.generate(prompt, cfg);
    return 0;
}
