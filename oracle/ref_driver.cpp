// ref_driver.cpp — thin command-line driver around the UNMODIFIED reference sources.
//
// TEST INFRASTRUCTURE ONLY.  This file contains no reference code; it #includes the reference's
// own headers and is linked with the reference's own src/api_engine.cpp, src/api_segment.cpp,
// src/api_autocomplete.cpp, src/api_metadata.cpp and src/semantic_embedding.cpp, compiled where
// they lie under /root/reference by oracle/Makefile into oracle/_ref/ref_engine (git-ignored).
// It gives tests/ and bench.py access to the real cord19::Engine::reload()/search() and the real
// SegmentWriter so that the oracle restatement and the CUDA path can be pinned against them.
//
// Modes
//   ref_engine write <dump> <segdir>
//        feed a corpus dump (written by ns_corpus_write_segment(dump_path=…)) through
//        SegmentWriter::add_document / write_segment (include/segment_writer.hpp:48-168)
//   ref_engine manifest <index_dir> <seg>...          save_manifest (src/api_segment.cpp:29-35)
//   ref_engine search <index_dir> <queries.txt> <k> <out.jsonl|-> [first] [count] [latencies.txt]
//        Engine::reload() + Engine::search(line, k) per line, run from a fresh empty CWD so that no
//        stale search_cache.json is served (src/api_engine.cpp:156,380-385); one JSON object per
//        query is written to out.jsonl ("-" = none), a timing summary to stdout.
#include <unistd.h>

#include <algorithm>
#include <chrono>
#include <cstdio>
#include <cstring>
#include <fstream>
#include <iostream>
#include <string>
#include <vector>

#include "api_engine.hpp"
#include "api_segment.hpp"
#include "segment_writer.hpp"
#include "textutil.hpp"

using cord19::json;

static uint32_t get_u32(std::ifstream& in) { uint32_t v = 0; in.read((char*)&v, 4); return v; }
static std::string get_str(std::ifstream& in) {
    uint32_t n = get_u32(in);
    std::string s(n, '\0');
    if (n) in.read(&s[0], n);
    return s;
}

static int mode_write(const char* dump, const char* segdir) {
    std::ifstream in(dump, std::ios::binary);
    if (!in) { std::cerr << "cannot open dump " << dump << "\n"; return 2; }
    SegmentWriter w;
    uint32_t ndocs = get_u32(in);
    for (uint32_t d = 0; d < ndocs; d++) {
        DocMeta m;
        m.cord_uid = get_str(in);
        m.doc_len = get_u32(in);
        uint32_t n = get_u32(in);
        std::vector<std::pair<std::string, uint32_t>> tfs;
        tfs.reserve(n);
        for (uint32_t i = 0; i < n; i++) {
            std::string t = get_str(in);
            uint32_t tf = get_u32(in);
            tfs.push_back({std::move(t), tf});
        }
        w.add_document(m, tfs);
    }
    if (!in) { std::cerr << "truncated dump\n"; return 2; }
    w.write_segment(segdir);
    return 0;
}

// query files are one query per line; "\\n" and "\\\\" escape a newline and a backslash
static std::string unescape(const std::string& s) {
    std::string o;
    for (size_t i = 0; i < s.size(); i++) {
        if (s[i] == '\\' && i + 1 < s.size()) {
            o.push_back(s[i + 1] == 'n' ? '\n' : s[i + 1]);
            i++;
        } else {
            o.push_back(s[i]);
        }
    }
    return o;
}

static int mode_search(int argc, char** argv) {
    std::string index_dir = fs::absolute(argv[2]).string();
    std::string qfile = fs::absolute(argv[3]).string();
    int k = std::atoi(argv[4]);
    std::string out = std::strcmp(argv[5], "-") == 0 ? std::string() : fs::absolute(argv[5]).string();
    long first = argc > 6 ? std::atol(argv[6]) : 0;
    long count = argc > 7 ? std::atol(argv[7]) : -1;
    std::string latfile = argc > 8 ? fs::absolute(argv[8]).string() : std::string();

    std::vector<std::string> queries;
    {
        std::ifstream in(qfile);
        std::string line;
        long i = 0;
        while (std::getline(in, line)) {
            if (i >= first && (count < 0 || i < first + count)) queries.push_back(unescape(line));
            i++;
        }
    }
    // fresh CWD: the engine reads and rewrites search_cache.json & co. in the working directory
    char tmpl[] = "/tmp/ref_engine_cwd_XXXXXX";
    char* cwd = mkdtemp(tmpl);
    if (!cwd || chdir(cwd) != 0) { std::cerr << "cannot create temp cwd\n"; return 2; }

    using clk = std::chrono::steady_clock;
    double reload_s = 0, search_s = 0;
    int rc = 0;
    {
        cord19::Engine engine;
        engine.index_dir = index_dir;
        auto t0 = clk::now();
        bool ok = engine.reload();
        reload_s = std::chrono::duration<double>(clk::now() - t0).count();
        if (!ok) { std::cerr << "reload failed\n"; rc = 3; }
        std::ofstream os;
        if (ok && !out.empty()) os.open(out);
        std::vector<double> lat;
        lat.reserve(queries.size());
        if (ok) {
            for (auto& q : queries) {
                auto s0 = clk::now();
                json j = engine.search(q, k);
                double dt = std::chrono::duration<double>(clk::now() - s0).count();
                search_s += dt;
                lat.push_back(dt);
                if (os.is_open()) {
                    // the exact text the reference's HTTP layer would send (res.set_content(j.dump(2)...) aside,
                    // dump() is the canonical form): kept verbatim for byte-level comparison
                    const std::string text = j.dump();
                    // qterms_w as Engine::search computed it (src/api_engine.cpp:410-417): expand() is const and
                    // deterministic, so calling it again with the same base terms yields the same list
                    if (engine.sem.enabled) {
                        std::vector<std::string> base;
                        for (auto& t : tokenize(q)) {
                            if (t.size() < 2) continue;
                            if (is_stopword(t)) continue;
                            base.push_back(t);
                        }
                        json qt = json::array();
                        if (!base.empty()) {
                            for (auto& tw : engine.sem.expand(base, 3, 5, 0.55f, 0.6f, 40)) {
                                uint32_t wb;
                                std::memcpy(&wb, &tw.second, 4);
                                qt.push_back(json::array({tw.first, wb}));
                            }
                        }
                        j["_qterms"] = qt;
                    }
                    j["_text"] = text;
                    // score as the exact f32 bit pattern: r["score"] holds the float widened to double
                    for (auto& r : j["results"]) {
                        float f = (float)r["score"].get<double>();
                        uint32_t bits;
                        std::memcpy(&bits, &f, 4);
                        r["score_bits"] = bits;
                    }
                    os << j.dump() << "\n";
                }
            }
        }
        if (!latfile.empty()) {
            std::ofstream lf(latfile);
            lf.precision(9);
            for (double v : lat) lf << v << "\n";
        }
        std::sort(lat.begin(), lat.end());
        json s;
        s["queries"] = queries.size();
        s["reload_seconds"] = reload_s;
        s["search_seconds"] = search_s;
        s["qps"] = search_s > 0 ? (double)queries.size() / search_s : 0.0;
        s["p50_ms"] = lat.empty() ? 0.0 : lat[lat.size() / 2] * 1e3;
        s["p99_ms"] = lat.empty() ? 0.0 : lat[(size_t)((double)(lat.size() - 1) * 0.99)] * 1e3;
        std::cout << s.dump() << std::endl;
    }  // ~Engine writes its caches into the temp cwd
    if (chdir("/") == 0) {
        std::error_code ec;
        fs::remove_all(cwd, ec);
    }
    return rc;
}

int main(int argc, char** argv) {
    if (argc >= 4 && std::strcmp(argv[1], "write") == 0) return mode_write(argv[2], argv[3]);
    if (argc >= 3 && std::strcmp(argv[1], "manifest") == 0) {
        std::vector<std::string> segs;
        for (int i = 3; i < argc; i++) segs.push_back(argv[i]);
        fs::create_directories(argv[2]);
        cord19::save_manifest(fs::path(argv[2]) / "manifest.bin", segs);
        return 0;
    }
    if (argc >= 6 && std::strcmp(argv[1], "search") == 0) return mode_search(argc, argv);
    std::cerr << "usage: ref_engine write <dump> <segdir> | manifest <index_dir> <seg>... | "
                 "search <index_dir> <queries.txt> <k> <out.jsonl|-> [first] [count] [latencies.txt]\n";
    return 64;
}
