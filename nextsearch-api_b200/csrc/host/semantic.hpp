// Semantic query expansion: the (term, qweight) list Engine::search scores when an
// embeddings file is present (src/api_engine.cpp:116-152, 410-421; src/semantic_embedding.cpp:34-229).
//
// What must match the reference bit for bit is the LIST — terms, f32 weights and their ORDER, because the
// order of qterms_w is the order in which a document's term scores are added (src/api_engine.cpp:449-480)
// and float addition does not associate.  The reference's order is an artefact of two libstdc++ facilities:
// iteration over a std::unordered_map<std::string, float> filled in a particular sequence, then a (non-stable)
// std::sort by weight.  The same containers, filled in the same sequence, are used here on purpose; everything
// around them — file parsing, the similarity scan, the neighbour selection — is written for this code base:
// the scan is a separable step (`Scanner`) so that the device kernel (cosine_scan_kernel, semantic_kernels.cuh)
// can stand in for the host loop; both produce the survivors (sim not below min_sim) in ascending row
// order with sims computed as the reference computes them (sequential f32 multiply-then-add, no FMA), and
// the reference's bounded-heap selection is then replayed over the survivors only.
#pragma once
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <string>
#include <unordered_map>
#include <unordered_set>
#include <utility>
#include <vector>

#include "common.hpp"

namespace nsb {

struct SemanticIndex {
    bool enabled = false;
    int dim = 0;
    std::vector<std::string> terms;                       // row -> term
    std::vector<float> vecs;                              // row-major, L2-normalised
    std::unordered_map<std::string, uint32_t> term_to_row;  // first row of a word wins

    struct Survivor {
        uint32_t row;
        float sim;
    };
    // survivors of M query vectors (qvecs[M][dim]): per vector, every row whose sim is not below min_sim, ascending
    // row.  Banned rows are dropped afterwards, on the host.
    using Scanner = std::function<void(const float* qvecs, uint32_t M, float min_sim, std::vector<std::vector<Survivor>>& out)>;
    Scanner device_scan;  // set by the engine when the vectors are resident on a GPU; empty = host loop

    // the reference's constants (src/api_engine.cpp:412-417)
    static constexpr int kPerTerm = 3, kGlobalTopk = 5, kMaxTotal = 40;
    static constexpr float kMinSim = 0.55f, kAlpha = 0.6f;

    static void l2_normalize(std::vector<float>& v) {  // src/semantic_embedding.cpp:18-24
        double ss = 0.0;
        for (float x : v) ss += (double)x * (double)x;
        const double n = std::sqrt(ss);
        if (n <= 0.0) return;
        for (float& x : v) x = (float)((double)x / n);
    }

    // Text format: one "word v1 ... vD" per line, optional "<count> <dim>" header (src/semantic_embedding.cpp:34-101).
    // `needed` filters words (the engine passes "is in some segment's lexicon"); an always-true filter loads all.
    bool load(const std::string& path, const std::function<bool(const std::string&)>& needed) {
        enabled = false;
        dim = 0;
        terms.clear();
        vecs.clear();
        term_to_row.clear();
        std::vector<uint8_t> bytes;
        if (!read_file(path, bytes)) return false;
        const char* p = (const char*)bytes.data();
        const char* end = p + bytes.size();
        bool first_line = true;
        size_t loaded = 0;
        std::vector<float> v;
        std::string word;
        while (p < end) {
            const char* nl = (const char*)std::memchr(p, '\n', (size_t)(end - p));
            const char* le = nl ? nl : end;
            const char* q = p;
            p = nl ? nl + 1 : end;
            if (le == q) continue;  // empty line
            if (first_line) {
                first_line = false;
                if (is_header(q, le)) continue;
            }
            q = skip_ws(q, le);
            const char* w0 = q;
            while (q < le && !is_ws(*q)) q++;
            if (q == w0) continue;
            word.assign(w0, q);
            if (!needed(word)) continue;
            v.clear();
            for (;;) {
                q = skip_ws(q, le);
                float x;
                const char* after = parse_float(q, le, x);
                if (!after) break;
                v.push_back(x);
                q = after;
            }
            if (v.size() < 10) continue;
            if (dim == 0) dim = (int)v.size();
            if ((int)v.size() != dim) continue;
            l2_normalize(v);
            const uint32_t row = (uint32_t)terms.size();
            terms.push_back(word);
            term_to_row.emplace(word, row);
            vecs.insert(vecs.end(), v.begin(), v.end());
            loaded++;
        }
        enabled = loaded > 0 && dim > 0;
        return enabled;
    }

    const float* vec_of(const std::string& term) const {
        auto it = term_to_row.find(term);
        return it == term_to_row.end() ? nullptr : &vecs[(size_t)it->second * (size_t)dim];
    }

    void host_scan_one(const float* qvec, float min_sim, std::vector<Survivor>& out) const {
        out.clear();
        const size_t nrows = terms.size();
        for (size_t r = 0; r < nrows; r++) {
            const float* v = &vecs[r * (size_t)dim];
            float s = 0.0f;
            for (int i = 0; i < dim; i++) s += qvec[i] * v[i];  // two roundings per step (the build uses -ffp-contract=off)
            if (s < min_sim) continue;
            out.push_back(Survivor{(uint32_t)r, s});
        }
    }
    void host_scan(const float* qvecs, uint32_t M, float min_sim, std::vector<std::vector<Survivor>>& out) const {
        out.resize(M);
        for (uint32_t m = 0; m < M; m++) host_scan_one(qvecs + (size_t)m * (size_t)dim, min_sim, out[m]);
    }

    // The reference keeps the best `topk` of the scan in a bounded min-heap and sorts it at the end
    // (most_similar_to_vec, src/semantic_embedding.cpp:104-146); replaying exactly that over the survivors
    // reproduces its result including the arbitrary choices among equal similarities.
    static void select_topk(const std::vector<Survivor>& surv, int topk, std::vector<Survivor>& out) {
        out.clear();
        if (topk <= 0) return;
        auto cmp = [](const Survivor& a, const Survivor& b) { return a.sim > b.sim; };
        std::vector<Survivor> heap;
        heap.reserve((size_t)topk);
        for (const Survivor& s : surv) {
            if ((int)heap.size() < topk) {
                heap.push_back(s);
                std::push_heap(heap.begin(), heap.end(), cmp);
            } else if (s.sim > heap.front().sim) {
                std::pop_heap(heap.begin(), heap.end(), cmp);
                heap.back() = s;
                std::push_heap(heap.begin(), heap.end(), cmp);
            }
        }
        std::sort_heap(heap.begin(), heap.end(), cmp);
        std::reverse(heap.begin(), heap.end());
        out = heap;
    }

    // qterms_w for a non-empty list of base terms (SemanticIndex::expand with the engine's constants)
    std::vector<std::pair<std::string, float>> expand(const std::vector<std::string>& base) const {
        std::unordered_map<std::string, float> w;
        w.reserve((size_t)kMaxTotal * 2);
        for (const auto& t : base)
            if (!t.empty()) w[t] = 1.0f;
        std::unordered_set<uint32_t> banned;
        banned.reserve(base.size() * 2);
        for (const auto& t : base) {
            auto it = term_to_row.find(t);
            if (it != term_to_row.end()) banned.insert(it->second);
        }
        auto offer = [&](const std::vector<Survivor>& nn, float cap) {
            for (const Survivor& s : nn) {
                const std::string& cand = terms[s.row];
                const float weight = std::max(0.0f, std::min(cap, cap * s.sim));
                auto it = w.find(cand);
                if (it == w.end() || weight > it->second) w[cand] = weight;
            }
        };
        // every vector this query scans with: one per base term that has a vector (in query order), then the
        // normalised centroid of those — ONE batched scan (the reference scans once per vector)
        std::vector<float> qbuf;
        uint32_t nper = 0;
        std::vector<float> centroid((size_t)dim, 0.0f);
        for (const auto& t : base) {
            const float* v = vec_of(t);
            if (!v) continue;
            qbuf.insert(qbuf.end(), v, v + dim);
            for (int j = 0; j < dim; j++) centroid[(size_t)j] += v[j];
            nper++;
        }
        if (nper > 0) {
            for (int j = 0; j < dim; j++) centroid[(size_t)j] /= (float)nper;
            l2_normalize(centroid);
            qbuf.insert(qbuf.end(), centroid.begin(), centroid.end());
        }
        std::vector<std::vector<Survivor>> surv;
        const uint32_t M = nper > 0 ? nper + 1 : 0;
        if (M) {
            if (device_scan) device_scan(qbuf.data(), M, kMinSim, surv);
            else host_scan(qbuf.data(), M, kMinSim, surv);
        }
        std::vector<Survivor> kept, best;
        auto neighbours = [&](uint32_t m, int topk, float cap) {
            kept.clear();
            for (const Survivor& sv : surv[m])
                if (banned.find(sv.row) == banned.end()) kept.push_back(sv);
            select_topk(kept, topk, best);
            offer(best, cap);
        };
        for (uint32_t m = 0; m < nper; m++) neighbours(m, kPerTerm, kAlpha);  // neighbours of every base term
        if (nper > 0) neighbours(nper, kGlobalTopk, kAlpha * 0.8f);           // neighbours of the centroid
        std::vector<std::pair<std::string, float>> out;
        out.reserve(w.size());
        for (auto& kv : w) out.push_back(kv);
        std::sort(out.begin(), out.end(), [](const auto& a, const auto& b) { return a.second > b.second; });
        if ((int)out.size() > kMaxTotal) out.resize((size_t)kMaxTotal);
        return out;
    }

  private:
    static bool is_ws(char c) { return c == ' ' || (c >= '\t' && c <= '\r'); }
    static const char* skip_ws(const char* p, const char* e) {
        while (p < e && is_ws(*p)) p++;
        return p;
    }
    // "<a> <b>" with nothing else, a > 0, 0 < b < 5000
    static bool is_header(const char* p, const char* e) {
        long long v[2];
        for (int i = 0; i < 2; i++) {
            p = skip_ws(p, e);
            const char* s = p;
            if (p < e && (*p == '+' || *p == '-')) p++;
            const char* d = p;
            while (p < e && *p >= '0' && *p <= '9') p++;
            if (p == d) return false;
            v[i] = std::strtoll(std::string(s, p).c_str(), nullptr, 10);
            if (p < e && !is_ws(*p)) return false;
        }
        p = skip_ws(p, e);
        return p == e && v[0] > 0 && v[1] > 0 && v[1] < 5000;
    }
    // one decimal floating-point token: [+-] digits [. digits] [e[+-]digits]; nullptr if none starts at p
    static const char* parse_float(const char* p, const char* e, float& out) {
        const char* s = p;
        if (p < e && (*p == '+' || *p == '-')) p++;
        const char* m = p;
        while (p < e && *p >= '0' && *p <= '9') p++;
        size_t digits = (size_t)(p - m);
        if (p < e && *p == '.') {
            p++;
            const char* f = p;
            while (p < e && *p >= '0' && *p <= '9') p++;
            digits += (size_t)(p - f);
        }
        if (digits == 0) return nullptr;
        if (p < e && (*p == 'e' || *p == 'E')) {
            const char* x = p + 1;
            if (x < e && (*x == '+' || *x == '-')) x++;
            const char* xd = x;
            while (x < e && *x >= '0' && *x <= '9') x++;
            if (x > xd) p = x;
        }
        out = std::strtof(std::string(s, p).c_str(), nullptr);
        return p;
    }
};

}  // namespace nsb
