// Sanitizer drivers for the HOST side of libnsb200 (no GPU needed: a host-only engine loads segments and resolves
// queries; it never scores).  Built by tests/test_host_sanitizers.py from the product's own host sources with
// -fsanitize=address,undefined (fuzzload) or -fsanitize=thread (stress).
//
//   host_stress fuzzload <index_dir> <scratch_dir> <iterations> <seed>
//       copies the index, damages ONE file of one segment (truncation, byte flips, a count field overwritten with a
//       huge value), reloads: every outcome but a crash / sanitizer report is fine (NS_OK or an error code + text).
//   host_stress stress <dirA> <dirB> <link> <threads> <seconds>
//       <link> is a symlink the engine was created on; a reloader thread flips it between two different corpora and
//       reloads while <threads> workers resolve batches and read names / stats / uids: every resolved batch must be
//       the answer of ONE corpus as a whole (generation snapshot), never a mixture.
#include <atomic>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <filesystem>
#include <fstream>
#include <random>
#include <string>
#include <thread>
#include <vector>

#include "../../include/nextsearch_b200.h"

namespace fs = std::filesystem;

static std::vector<char> slurp(const fs::path& p) {
    std::ifstream in(p, std::ios::binary);
    return std::vector<char>((std::istreambuf_iterator<char>(in)), std::istreambuf_iterator<char>());
}
static void spit(const fs::path& p, const std::vector<char>& b) {
    std::ofstream out(p, std::ios::binary | std::ios::trunc);
    out.write(b.data(), (std::streamsize)b.size());
}

static int fuzzload(const char* index_dir, const char* scratch, int iters, unsigned seed) {
    std::mt19937 rng(seed);
    int ok = 0, refused = 0;
    std::vector<fs::path> files;
    for (auto& e : fs::recursive_directory_iterator(index_dir))
        if (e.is_regular_file()) files.push_back(fs::relative(e.path(), index_dir));
    if (files.empty()) { std::fprintf(stderr, "no files under %s\n", index_dir); return 2; }
    for (int it = 0; it < iters; it++) {
        fs::path work = fs::path(scratch) / ("fz" + std::to_string(it % 4));
        fs::remove_all(work);
        fs::copy(index_dir, work, fs::copy_options::recursive);
        // prefer the small structured files; inverted barrels are validated on the device, not here.  Three times in ten
        // the victim is a top-level file: manifest.bin, metadata.csv, the embeddings file.
        std::vector<fs::path> top;
        for (auto& f : files)
            if (!f.has_parent_path()) top.push_back(f);
        fs::path victim;
        for (int tries = 0; tries < 64; tries++) {
            victim = (!top.empty() && rng() % 10 < 3) ? top[rng() % top.size()] : files[rng() % files.size()];
            const std::string n = victim.filename().string();
            if (n.rfind("inverted", 0) != 0 && n.rfind("forward", 0) != 0 && n.rfind("terms", 0) != 0) break;
        }
        std::vector<char> b = slurp(work / victim);
        const int how = (int)(rng() % 4);
        if (how == 0) {
            b.resize(b.empty() ? 0 : rng() % b.size());                                  // truncate
        } else if (how == 1 && !b.empty()) {
            for (int k = 0; k < 1 + (int)(rng() % 8); k++) b[rng() % b.size()] = (char)rng();   // byte flips
        } else if (how == 2 && b.size() >= 4) {
            const uint32_t huge[4] = {0xFFFFFFFFu, 0x7FFFFFFFu, 0x10000000u, 0x00FFFFFFu};
            const size_t at = (rng() % 3 == 0 || b.size() < 8) ? 0 : (rng() % (b.size() - 3));
            std::memcpy(b.data() + at, &huge[rng() % 4], 4);                             // a count / length field blown up
        } else {
            b.clear();                                                                   // empty file
        }
        spit(work / victim, b);
        ns_engine* e = nullptr;
        if (ns_engine_create(work.string().c_str(), -1, &e) != NS_OK) { std::fprintf(stderr, "create failed\n"); return 2; }
        const int rc = ns_engine_reload(e);
        if (rc == NS_OK) {
            ok++;
            // a reload that succeeded must leave a usable front end
            const char* qs[2] = {"t1 t2 t3", "alpha beta"};
            uint64_t q_off[3], n_terms = 0;
            std::vector<ns_qterm> terms(4096);
            uint8_t has[2];
            if (ns_engine_resolve_batch(e, 2, qs, q_off, terms.data(), terms.size(), &n_terms, has) != NS_OK) {
                std::fprintf(stderr, "resolve after a successful reload failed: %s\n", ns_last_error());
                return 3;
            }
        } else {
            refused++;
            if (!ns_last_error() || !*ns_last_error()) { std::fprintf(stderr, "refusal without a text (rc %d, %s)\n", rc, victim.string().c_str()); return 3; }
        }
        ns_engine_destroy(e);
    }
    std::printf("{\"mode\": \"fuzzload\", \"iterations\": %d, \"loaded\": %d, \"refused\": %d}\n", iters, ok, refused);
    return 0;
}

static int stress(const char* dirA, const char* dirB, const char* link, int nthreads, double seconds) {
    fs::remove(link);
    fs::create_directory_symlink(fs::absolute(dirA), link);
    ns_engine* e = nullptr;
    if (ns_engine_create(link, -1, &e) != NS_OK || ns_engine_reload(e) != NS_OK) { std::fprintf(stderr, "setup: %s\n", ns_last_error()); return 2; }
    // the batch and its two legitimate answers
    std::vector<std::string> qs;
    for (int i = 0; i < 300; i++) qs.push_back("t" + std::to_string(1 + i % 97) + " t" + std::to_string(1 + (i * 7) % 211) + " the t" + std::to_string(3 + i));
    std::string packed;
    for (auto& q : qs) { packed += q; packed.push_back('\0'); }
    const uint32_t Q = (uint32_t)qs.size();
    auto answer = [&](std::vector<uint64_t>& q_off, std::vector<ns_qterm>& terms) -> int {
        q_off.assign(Q + 1, 0);
        uint64_t n = 0;
        terms.assign(1 << 16, ns_qterm{});
        std::vector<uint8_t> has(Q);
        int rc = ns_engine_resolve_batch_packed(e, Q, packed.data(), packed.size(), q_off.data(), terms.data(), terms.size(), &n, has.data());
        terms.resize(n);
        return rc;
    };
    std::vector<uint64_t> offA, offB;
    std::vector<ns_qterm> tA, tB;
    if (answer(offA, tA) != NS_OK) return 2;
    fs::remove(link);
    fs::create_directory_symlink(fs::absolute(dirB), link);
    if (ns_engine_reload(e) != NS_OK || answer(offB, tB) != NS_OK) return 2;
    auto same = [](const std::vector<ns_qterm>& x, const std::vector<ns_qterm>& y) {
        return x.size() == y.size() && (x.empty() || std::memcmp(x.data(), y.data(), x.size() * sizeof(ns_qterm)) == 0);
    };
    if (same(tA, tB)) { std::fprintf(stderr, "the two corpora resolve identically: the test would prove nothing\n"); return 2; }
    std::atomic<bool> stop{false};
    std::atomic<long> batches{0}, mixed{0}, errors{0}, reloads{0};
    std::vector<std::thread> th;
    for (int t = 0; t < nthreads; t++)
        th.emplace_back([&, t] {
            std::vector<uint64_t> off;
            std::vector<ns_qterm> tt;
            char buf[256];
            while (!stop.load()) {
                if (answer(off, tt) != NS_OK) { errors++; continue; }
                if (!(same(tt, tA) && off == offA) && !(same(tt, tB) && off == offB)) mixed++;
                batches++;
                if (t % 2 == 0) {   // the small read-only entry points race with the swap as well
                    (void)ns_engine_num_segments(e);
                    (void)ns_engine_segment_name(e, 0, buf, sizeof buf);
                    uint32_t df = 0, cnt = 0;
                    (void)ns_engine_term_stats(e, 0, "t1", &df, &cnt);
                    (void)ns_engine_cord_uid(e, 0, 3, buf, sizeof buf);
                    uint32_t N = 0, T = 0;
                    float avg = 0;
                    uint64_t P = 0;
                    (void)ns_engine_segment_stats(e, 0, &N, &avg, &T, &P);
                }
            }
        });
    std::thread reloader([&] {
        bool toA = true;
        while (!stop.load()) {
            fs::remove(link);
            fs::create_directory_symlink(fs::absolute(toA ? dirA : dirB), link);
            if (ns_engine_reload(e) != NS_OK) errors++;
            reloads++;
            toA = !toA;
            std::this_thread::sleep_for(std::chrono::milliseconds(2));
        }
    });
    std::this_thread::sleep_for(std::chrono::milliseconds((long)(seconds * 1000)));
    stop = true;
    for (auto& x : th) x.join();
    reloader.join();
    ns_engine_destroy(e);
    std::printf("{\"mode\": \"stress\", \"threads\": %d, \"batches\": %ld, \"reloads\": %ld, \"mixed\": %ld, \"errors\": %ld}\n", nthreads,
                batches.load(), reloads.load(), mixed.load(), errors.load());
    return (mixed.load() || errors.load()) ? 4 : 0;
}

int main(int argc, char** argv) {
    if (argc >= 6 && !std::strcmp(argv[1], "fuzzload")) return fuzzload(argv[2], argv[3], std::atoi(argv[4]), (unsigned)std::atoi(argv[5]));
    if (argc >= 7 && !std::strcmp(argv[1], "stress")) return stress(argv[2], argv[3], argv[4], std::atoi(argv[5]), std::atof(argv[6]));
    std::fprintf(stderr, "usage: host_stress fuzzload <index_dir> <scratch_dir> <iterations> <seed> | stress <dirA> <dirB> <link> <threads> <seconds>\n");
    return 64;
}
