#include "text.h"
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <limits>

namespace nt { namespace b200 {

namespace {
// GPT-2 byte <-> unicode table: printable bytes map to themselves, the rest to U+0100.. in order.
struct ByteMap {
    std::string enc[256];
    std::unordered_map<std::string, uint8_t> dec;
    ByteMap() {
        int extra = 0;
        for (int b = 0; b < 256; b++) {
            bool keep = (b >= 33 && b <= 126) || (b >= 161 && b <= 172) || (b >= 174);
            uint32_t cp = keep ? (uint32_t)b : 256u + (uint32_t)extra++;
            std::string u;
            if (cp < 0x80) u += (char)cp;
            else if (cp < 0x800) { u += (char)(0xC0 | (cp >> 6)); u += (char)(0x80 | (cp & 0x3F)); }
            else { u += (char)(0xE0 | (cp >> 12)); u += (char)(0x80 | ((cp >> 6) & 0x3F)); u += (char)(0x80 | (cp & 0x3F)); }
            enc[b] = u;
            dec[u] = (uint8_t)b;
        }
    }
};
const ByteMap& bytemap() { static ByteMap m; return m; }
int utf8_len(uint8_t c) { return c < 0x80 ? 1 : (c & 0xE0) == 0xC0 ? 2 : (c & 0xF0) == 0xE0 ? 3 : (c & 0xF8) == 0xF0 ? 4 : 0; }
}  // namespace

void Tokenizer::init(const GGUFVocab& v, int bos_id, int eos_id) {
    tokens_ = v.tokens; scores_ = v.scores; types_ = v.token_types;
    bos_ = bos_id; eos_ = eos_id;
    ids_.clear();
    ids_.reserve(tokens_.size() * 2);
    for (int i = 0; i < (int)tokens_.size(); i++) ids_[tokens_[i]] = i;       // later duplicates win, like the reference
    gpt2_ = ids_.count(bytemap().enc[0x20]) > 0;
    fprintf(stderr, "Tokenizer: %d tokens, BOS=%d, EOS=%d, encoding=%s\n", (int)tokens_.size(), bos_, eos_,
            gpt2_ ? "GPT2-BPE" : "SentencePiece");
}

int Tokenizer::byte_token(uint8_t b) const {
    if (gpt2_) { auto it = ids_.find(bytemap().enc[b]); if (it != ids_.end()) return it->second; }
    char name[8];
    snprintf(name, sizeof(name), "<0x%02X>", b);
    auto it = ids_.find(name);
    return it == ids_.end() ? 0 : it->second;
}

std::vector<int> Tokenizer::encode(const std::string& text, bool add_bos) const {
    std::vector<int> out;
    if (add_bos) out.push_back(bos_);
    if (text.empty()) return out;
    std::string enc;
    if (gpt2_) { for (unsigned char c : text) enc += bytemap().enc[c]; }
    else { for (char c : text) { if (c == ' ') enc += "\xe2\x96\x81"; else enc += c; } }

    struct Sym { int id; std::string text; int next; };
    std::vector<Sym> syms;
    for (size_t pos = 0; pos < enc.size();) {                 // greedy longest match, window 64 bytes
        size_t best = 0; int id = -1;
        for (size_t len = std::min<size_t>(64, enc.size() - pos); len >= 1; len--) {
            auto it = ids_.find(enc.substr(pos, len));
            if (it != ids_.end()) { best = len; id = it->second; break; }
        }
        if (!best) { best = 1; id = byte_token((uint8_t)enc[pos]); }
        if (!syms.empty()) syms.back().next = (int)syms.size();
        syms.push_back({id, enc.substr(pos, best), -1});
        pos += best;
    }
    for (;;) {                                                // merge the adjacent pair whose merged token scores highest
        float best_score = -std::numeric_limits<float>::infinity();
        int bi = -1;
        for (int i = 0; i < (int)syms.size(); i++) {
            if (syms[i].next < 0 || syms[i].id < 0) continue;
            auto it = ids_.find(syms[i].text + syms[syms[i].next].text);
            if (it == ids_.end()) continue;
            float sc = it->second < (int)scores_.size() ? scores_[it->second] : 0.0f;
            if (sc > best_score) { best_score = sc; bi = i; }
        }
        if (bi < 0) break;
        int j = syms[bi].next;
        syms[bi].text += syms[j].text;
        syms[bi].id = ids_.find(syms[bi].text)->second;
        syms[bi].next = syms[j].next;
        syms[j].id = -1;
    }
    for (const Sym& s : syms) if (s.id >= 0) out.push_back(s.id);
    return out;
}

std::string Tokenizer::decode_token(int id) const {
    if (id < 0 || id >= (int)tokens_.size()) return "";
    if (id < (int)types_.size() && (types_[id] == 3 || types_[id] == 4)) return "";   // control / unused
    const std::string& t = tokens_[id];
    std::string r;
    if (gpt2_) {
        for (size_t pos = 0; pos < t.size();) {
            int n = utf8_len((uint8_t)t[pos]);
            if (!n) { pos++; continue; }
            if (pos + n > t.size()) break;
            std::string ch = t.substr(pos, n);
            auto it = bytemap().dec.find(ch);
            if (it != bytemap().dec.end()) r += (char)it->second; else r += ch;
            pos += n;
        }
        return r;
    }
    if (t.size() == 6 && t[0] == '<' && t[1] == '0' && t[2] == 'x' && t[5] == '>') {
        char hex[3] = {t[3], t[4], 0};
        return std::string(1, (char)strtol(hex, nullptr, 16));
    }
    for (size_t pos = 0; pos < t.size();) {
        if (pos + 2 < t.size() && (uint8_t)t[pos] == 0xE2 && (uint8_t)t[pos + 1] == 0x96 && (uint8_t)t[pos + 2] == 0x81) { r += ' '; pos += 3; }
        else r += t[pos++];
    }
    return r;
}

std::string Tokenizer::decode(const std::vector<int>& ids) const {
    std::string r;
    for (int id : ids) r += decode_token(id);
    return r;
}

int Sampler::argmax(const float* logits, int n) {
    int best = 0;
    for (int i = 1; i < n; i++) if (logits[i] > logits[best]) best = i;
    return best;
}

void Sampler::apply_repeat_penalty(float* logits, int n, const std::vector<int>& recent) const {
    if (cfg_.repeat_penalty <= 1.0f) return;
    int window = std::min((int)recent.size(), cfg_.repeat_window);
    for (int i = (int)recent.size() - window; i < (int)recent.size(); i++) {
        int t = recent[i];
        if (t < 0 || t >= n) continue;
        if (logits[t] > 0) logits[t] /= cfg_.repeat_penalty; else logits[t] *= cfg_.repeat_penalty;
    }
}

int Sampler::sample(const float* logits, int n) {
    if (cfg_.temperature <= 0.0f) return argmax(logits, n);
    cand_.resize((size_t)n);
    for (int i = 0; i < n; i++) cand_[(size_t)i] = {logits[i] / cfg_.temperature, i};
    auto gt = [](const std::pair<float, int>& a, const std::pair<float, int>& b) { return a.first > b.first; };
    if (cfg_.top_k > 0 && cfg_.top_k < n) {
        std::partial_sort(cand_.begin(), cand_.begin() + cfg_.top_k, cand_.end(), gt);
        cand_.resize((size_t)cfg_.top_k);
    } else {
        std::sort(cand_.begin(), cand_.end(), gt);
    }
    float mx = cand_[0].first, sum = 0.f;
    for (auto& c : cand_) { c.first = expf(c.first - mx); sum += c.first; }
    for (auto& c : cand_) c.first /= sum;
    if (cfg_.top_p < 1.0f && cfg_.top_p > 0.0f) {
        float cum = 0.f;
        size_t cut = cand_.size();
        for (size_t i = 0; i < cand_.size(); i++) { cum += cand_[i].first; if (cum >= cfg_.top_p) { cut = i + 1; break; } }
        cand_.resize(cut);
        sum = 0.f;
        for (auto& c : cand_) sum += c.first;
        for (auto& c : cand_) c.first /= sum;
    }
    std::uniform_real_distribution<float> dist(0.0f, 1.0f);
    float r = dist(rng_), cum = 0.f;
    for (const auto& c : cand_) { cum += c.first; if (r <= cum) return c.second; }
    return cand_.back().second;
}

}}  // namespace nt::b200
