// main.cpp — CLI with the reference's flag surface (src/main.cpp:6-174).  The streaming family of flags
// (--streaming, --draft-model, --draft-k, --self-spec, --early-exit, --skip-threshold, --requant-q4k,
// --delta-model) belongs to the PCIe/NVMe streaming subsystem that this resident engine drops: they are
// rejected with an explanation instead of being silently ignored.
#include "engine.h"
#include <atomic>
#include <cstring>
#include <string>
#include <vector>
#include <sys/mman.h>
#include <sys/wait.h>
#include <unistd.h>

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
    fprintf(stderr, "  --device <int>           First CUDA device to use (default: 0)\n");
    fprintf(stderr, "  --tp <int>               Tensor-parallel over this many GPUs (devices device..device+tp-1, one process each,\n"
                    "                           attention heads and FFN columns sharded, partial sums exchanged over NVLink); 1, 2, 4 or 8\n");
    fprintf(stderr, "  --bpe-merges             Tokenise with the GGUF's rank-ordered BPE merges + literal special tokens (default: the reference's\n"
                    "                           score-based merging, which ignores tokenizer.ggml.merges)\n");
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
    int max_context = 4096, tp = 1, device = 0;
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
        else if (a == "--tp") { if (auto v = val()) tp = std::stoi(v); }
        else if (a == "--device") { if (auto v = val()) device = std::stoi(v); }
        else if (a == "--gpu-sampler") cfg.gpu_sampler = true;
        else if (a == "--host-sampler") cfg.gpu_sampler = false;
        else if (a == "--bpe-merges") setenv("NT_B200_BPE_MERGES", "1", 1);
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
    if (tp < 1 || tp > 8 || (tp & (tp - 1))) { fprintf(stderr, "Error: --tp must be 1, 2, 4 or 8\n"); return 1; }
    // Tensor parallel: one process per GPU.  The ranks are forked here, BEFORE anything touches CUDA; rank 0 hands the NCCL id to
    // the others through a shared page.  All ranks run the same generation with the same seed (bit-identical logits everywhere).
    struct Shared { std::atomic<int> ready; unsigned char id[TPComm::kIdBytes]; };
    Shared* sh = nullptr;
    int rank = 0;
    std::vector<pid_t> kids;
    if (tp > 1) {
        if (chat || (prompt.empty() && !bench)) { fprintf(stderr, "Error: --tp needs -p <prompt> or --benchmark (the ranks cannot share an interactive stdin)\n"); return 1; }
        sh = static_cast<Shared*>(mmap(nullptr, sizeof(Shared), PROT_READ | PROT_WRITE, MAP_SHARED | MAP_ANONYMOUS, -1, 0));
        if (sh == MAP_FAILED) { perror("mmap"); return 1; }
        new (&sh->ready) std::atomic<int>(0);
        for (int r = 1; r < tp; r++) {
            pid_t pid = fork();
            if (pid < 0) { perror("fork"); return 1; }
            if (pid == 0) { rank = r; kids.clear(); break; }
            kids.push_back(pid);
        }
    }
    if (cudaSetDevice(device + rank) != cudaSuccess) { fprintf(stderr, "Error: cannot select CUDA device %d\n", device + rank); return 1; }
    int rc = 0;
    {
        Engine engine;
        bool ok;
        if (tp > 1) {
            if (rank == 0) {
                if (!TPComm::unique_id(sh->id)) { fprintf(stderr, "Error: ncclGetUniqueId failed\n"); sh->ready.store(-1); return 1; }
                sh->ready.store(1);
            } else {
                while (sh->ready.load() == 0) usleep(200);
                if (sh->ready.load() < 0) return 1;
                cfg.verbose = false;                                  // only rank 0 talks
                if (!freopen("/dev/null", "w", stdout)) return 1;
                if (!getenv("NT_B200_TP_VERBOSE") && !freopen("/dev/null", "w", stderr)) return 1;   // loader / stats lines once, not tp times
            }
            ok = engine.load_tp(model_path, max_context, rank, tp, sh->id);
        } else {
            ok = engine.load(model_path, max_context);
        }
        if (!ok) { fprintf(stderr, "Failed to load model: %s\n", model_path.c_str()); rc = 1; }
        else if (bench) engine.benchmark(prompt.empty() ? "The meaning of life is" : prompt, cfg.max_tokens);
        else if (chat || prompt.empty()) engine.chat(cfg);
        else engine.generate(prompt, cfg);
    }
    for (pid_t k : kids) { int st = 0; waitpid(k, &st, 0); if (!WIFEXITED(st) || WEXITSTATUS(st) != 0) rc = 1; }
    return rc;
}
