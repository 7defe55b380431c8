// shim_engine.cpp — the reference-side binding of INTEGRATION.md §1, COMPILED against the reference's own
// include/api_engine.hpp (TEST INFRASTRUCTURE; built by oracle/Makefile into oracle/_ref/shim_engine).
//
// It defines the members of cord19::Engine that src/api_engine.cpp defines, with reload() and search()
// delegating to libnsb200.so through the C ABI (include/nextsearch_b200.h) — i.e. this file stands in for
// src/api_engine.cpp in a build of the reference's api_server — plus a small main() with the same `search`
// command line as oracle/ref_driver.cpp, so that tests can run the reference's own golden queries through
//     reference header  ->  this shim  ->  C ABI  ->  CUDA
// and compare the returned nlohmann::json (dumped with the reference's json library) with what the unmodified
// reference produced (tests/golden/*.json, field `text`).
//
// What the shim keeps from the reference's Engine and what it drops is stated in INTEGRATION.md: the LRU
// result cache and the AI caches are front-of-engine features outside the hot path; their members exist in the
// struct (the header is unmodified) and stay empty here.
#include <cstring>
#include <fstream>
#include <iostream>
#include <mutex>
#include <string>
#include <unordered_map>
#include <vector>

#include "api_engine.hpp"
#include "nextsearch_b200.h"

namespace cord19 {

namespace {
// the unmodified header has no room for the handle: keep it beside the object
std::mutex g_mu;
std::unordered_map<const Engine*, ns_engine*> g_handles;

ns_engine* handle_of(const Engine* e) {
    std::lock_guard<std::mutex> lk(g_mu);
    auto it = g_handles.find(e);
    return it == g_handles.end() ? nullptr : it->second;
}
}  // namespace

Engine::~Engine() {
    ns_engine* h = nullptr;
    {
        std::lock_guard<std::mutex> lk(g_mu);
        auto it = g_handles.find(this);
        if (it != g_handles.end()) {
            h = it->second;
            g_handles.erase(it);
        }
    }
    if (h) ns_engine_destroy(h);
}

// Engine::reload (src/api_engine.cpp:50-162): segments, lexicons, metadata and embeddings are loaded by
// ns_engine_reload; seg_names is filled for callers that read it (api_server prints the segment count).
bool Engine::reload() {
    std::lock_guard<std::mutex> lock(mtx);
    ns_engine* h = handle_of(this);
    if (!h) {
        int ndev = ns_device_count();
        if (const char* s = std::getenv("SHIM_DEVICES")) ndev = std::max(1, std::min(ndev, std::atoi(s)));
        std::vector<int> devs;
        for (int d = 0; d < ndev; d++) devs.push_back(d);
        if (ns_engine_create_multi(index_dir.string().c_str(), ndev, devs.data(), &h) != NS_OK || ndev == 0) {
            std::cerr << "[gpu] " << ns_last_error() << "\n";
            return false;  // no CPU fallback: fail the way a missing segment file does (:82-85)
        }
        std::lock_guard<std::mutex> lk(g_mu);
        g_handles[this] = h;
    }
    if (ns_engine_reload(h) != NS_OK) {  // the previous generation stays live on failure
        std::cerr << "[gpu] " << ns_last_error() << "\n";
        return false;
    }
    seg_names.clear();
    char name[512];
    for (int i = 0; i < ns_engine_num_segments(h); i++)
        if (ns_engine_segment_name(h, i, name, sizeof(name)) >= 0) seg_names.push_back(name);
    return true;
}

// Engine::search (src/api_engine.cpp:369-542).  No Engine::mtx around the GPU call: ns_engine_search_json is
// thread-safe, and with ns_engine_coalescer_start concurrent requests share one GPU batch.
json Engine::search(const std::string& query, int k) {
    ns_engine* h = handle_of(this);
    if (!h) throw std::runtime_error("search before reload");
    std::string buf(1 << 16, '\0');
    size_t need = 0;
    for (int attempt = 0; attempt < 2; attempt++) {
        if (ns_engine_search_json(h, query.c_str(), k, &buf[0], buf.size(), &need) != NS_OK)
            throw std::runtime_error(ns_last_error());  // -> HTTP 500 through the handler (src/api_server.cpp:76-84)
        if (need < buf.size()) break;
        buf.assign(need + 1, '\0');
    }
    return json::parse(buf.c_str());
}

// the rest of the struct's out-of-line members (front-of-engine features, not on the hot path)
json Engine::suggest(const std::string&, int) { return json::object(); }
std::string Engine::make_cache_key(const std::string& query, int k) { return query + "|" + std::to_string(k); }
json Engine::get_ai_overview_from_cache(const std::string&) { return json(); }
void Engine::put_ai_overview_in_cache(const std::string&, const json&) {}
json Engine::get_ai_summary_from_cache(const std::string&) { return json(); }
void Engine::put_ai_summary_in_cache(const std::string&, const json&) {}
void Engine::save_cache() {}
void Engine::load_cache() {}
void Engine::save_ai_overview_cache() {}
void Engine::load_ai_overview_cache() {}
void Engine::save_ai_summary_cache() {}
void Engine::load_ai_summary_cache() {}
json Engine::get_from_cache(const std::string&) { return json(); }
void Engine::put_in_cache(const std::string&, const json&) {}

}  // namespace cord19

using cord19::json;

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

// shim_engine search <index_dir> <queries.txt> <k> <out.jsonl>
int main(int argc, char** argv) {
    if (argc < 6 || std::strcmp(argv[1], "search") != 0) {
        std::cerr << "usage: shim_engine search <index_dir> <queries.txt> <k> <out.jsonl>\n";
        return 64;
    }
    cord19::Engine engine;
    engine.index_dir = cord19::fs::absolute(argv[2]);
    if (!engine.reload()) return 3;
    std::ifstream in(argv[3]);
    std::ofstream os(argv[5]);
    const int k = std::atoi(argv[4]);
    std::string line;
    while (std::getline(in, line)) {
        json j = engine.search(unescape(line), k);
        json row;
        row["text"] = j.dump();  // serialised by the reference's own json library
        os << row.dump() << "\n";
    }
    return 0;
}
