#include "text.h"
#include <algorithm>
#include <cmath>
#include <cctype>
#include <cstdio>
#include <cstdlib>
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
    merge_rank_.clear();
    merge_rank_.reserve(v.merges.size() * 2);
    for (int r = 0; r < (int)v.merges.size(); r++) {
        const std::string& m = v.merges[(size_t)r];
        const size_t sp = m.find(' ', 1);                    // the left symbol may itself start with the encoded space, never a raw ' '
        if (sp == std::string::npos || sp + 1 >= m.size()) continue;
        merge_rank_.emplace(m.substr(0, sp) + '\x01' + m.substr(sp + 1), r);
    }
    specials_.clear();
    for (int i = 0; i < (int)tokens_.size(); i++)
        if (i < (int)types_.size() && types_[i] == 3 && tokens_[i].size() >= 3 && tokens_[i].front() == '<' && tokens_[i].back() == '>')
            specials_.emplace_back(tokens_[i], i);
    std::sort(specials_.begin(), specials_.end(), [](const auto& a, const auto& b) { return a.first.size() > b.first.size(); });
    if (const char* e = getenv("NT_B200_BPE_MERGES")) use_merges_ = e[0] && !(e[0] == '0' && e[1] == 0);
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
    if (merges_active()) { encode_merges(text, out); return out; }
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

// ---- opt-in: rank-ordered byte-level BPE (tokenizer.ggml.merges) ----------------------------------------------------------------
namespace {
bool is_letter(unsigned char c) { return (c >= 'a' && c <= 'z') || (c >= 'A' && c <= 'Z') || c >= 0x80; }   // non-ASCII bytes count as letters
bool is_digit(unsigned char c) { return c >= '0' && c <= '9'; }
bool is_space(unsigned char c) { return c == ' ' || c == '\t' || c == '\n' || c == '\r' || c == '\f' || c == '\v'; }
bool is_newline(unsigned char c) { return c == '\n' || c == '\r'; }

// Llama-3 pre-tokeniser, byte classes in place of \p{L} / \p{N}:
//   (?i:'s|'t|'re|'ve|'m|'ll|'d) | [^\r\n L N]? L+ | N{1,3} | ' '? [^\s L N]+ [\r\n]* | \s* [\r\n]+ | \s+ (?!\S) | \s+
size_t next_piece(const std::string& t, size_t i) {
    const size_t n = t.size();
    auto at = [&](size_t k) -> unsigned char { return k < n ? (unsigned char)t[k] : 0; };
    if (t[i] == '\'' && i + 1 < n) {                                   // contractions
        const char a = (char)std::tolower(at(i + 1)), b = (char)std::tolower(at(i + 2));
        if (a == 's' || a == 't' || a == 'm' || a == 'd') return i + 2;
        if ((a == 'r' && b == 'e') || (a == 'v' && b == 'e') || (a == 'l' && b == 'l')) return i + 3;
    }
    {                                                                    // [^\r\n L N]? L+
        size_t j = i;
        if (!is_newline(at(j)) && !is_letter(at(j)) && !is_digit(at(j)) && j + 1 < n && is_letter(at(j + 1))) j++;
        if (j < n && is_letter(at(j))) { while (j < n && is_letter(at(j))) j++; return j; }
    }
    if (is_digit(at(i))) { size_t j = i; while (j < n && j < i + 3 && is_digit(at(j))) j++; return j; }
    {                                                                    // ' '? [^\s L N]+ [\r\n]*
        size_t j = i;
        if (at(j) == ' ') j++;
        if (j < n && !is_space(at(j)) && !is_letter(at(j)) && !is_digit(at(j))) {
            while (j < n && !is_space(at(j)) && !is_letter(at(j)) && !is_digit(at(j))) j++;
            while (j < n && is_newline(at(j))) j++;
            return j;
        }
    }
    if (is_space(at(i))) {
        size_t j = i, last_nl = 0;
        while (j < n && is_space(at(j))) { if (is_newline(at(j))) last_nl = j + 1; j++; }
        if (last_nl) return last_nl;                                     // \s* [\r\n]+
        if (j < n && j - i > 1) return j - 1;                            // \s+ (?!\S): leave the last space to the next word
        return j;
    }
    return i + 1;
}
}  // namespace

// One pre-token: its bytes as byte-level symbols, then the adjacent pair of lowest rank is merged until none is left.
void Tokenizer::bpe_word(const std::string& word, std::vector<int>& out) const {
    std::vector<std::string> sym;
    sym.reserve(word.size());
    for (unsigned char c : word) sym.push_back(bytemap().enc[c]);
    while (sym.size() > 1) {
        int best = -1, best_rank = std::numeric_limits<int>::max();
        for (int i = 0; i + 1 < (int)sym.size(); i++) {
            auto it = merge_rank_.find(sym[(size_t)i] + '\x01' + sym[(size_t)i + 1]);
            if (it != merge_rank_.end() && it->second < best_rank) { best_rank = it->second; best = i; }
        }
        if (best < 0) break;
        sym[(size_t)best] += sym[(size_t)best + 1];
        sym.erase(sym.begin() + best + 1);
    }
    for (const std::string& s : sym) {
        auto it = ids_.find(s);
        if (it != ids_.end()) { out.push_back(it->second); continue; }
        // a merged symbol that is not in the vocabulary (inconsistent merges): fall back to its bytes
        for (size_t pos = 0; pos < s.size();) {
            int n = utf8_len((uint8_t)s[pos]);
            if (!n || pos + n > s.size()) n = 1;
            auto d = bytemap().dec.find(s.substr(pos, (size_t)n));
            out.push_back(byte_token(d != bytemap().dec.end() ? d->second : (uint8_t)s[pos]));
            pos += (size_t)n;
        }
    }
}

void Tokenizer::encode_merges(const std::string& text, std::vector<int>& out) const {
    size_t i = 0, run = 0;                                   // [run, i) = ordinary text not yet tokenised
    auto flush = [&](size_t end) {
        const std::string part = text.substr(run, end - run);
        for (size_t p = 0; p < part.size();) { const size_t q = next_piece(part, p); bpe_word(part.substr(p, q - p), out); p = q; }
    };
    while (i < text.size()) {
        int hit = -1;
        if (text[i] == '<')
            for (int k = 0; k < (int)specials_.size(); k++)
                if (text.compare(i, specials_[(size_t)k].first.size(), specials_[(size_t)k].first) == 0) { hit = k; break; }
        if (hit < 0) { i++; continue; }
        flush(i);
        out.push_back(specials_[(size_t)hit].second);
        i += specials_[(size_t)hit].first.size();
        run = i;
    }
    flush(text.size());
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
