#include "segment_io.hpp"

#include <dirent.h>

#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstdio>
#include <thread>

#include "common.hpp"

namespace nsb {

float bm25_idf(uint32_t N, uint32_t df) {
    // (N - df) wraps in u32 exactly as in the reference when df > N.
    uint32_t nd = N - df;
    float num = (float)nd + 0.5f;
    float den = (float)df + 0.5f;
    float x = (num / den) + 1.0f;
    return std::log(x);  // float overload -> logf
}

std::string barrel_suffix(uint32_t b) {
    char buf[16];
    std::snprintf(buf, sizeof(buf), "%03u", b);
    return std::string(buf);
}

std::string seg_name(uint32_t id) {
    char buf[32];
    std::snprintf(buf, sizeof(buf), "seg_%06u", id);
    return std::string(buf);
}

std::vector<std::string> load_manifest(const std::string& manifest_path) {
    std::vector<std::string> segs;
    std::vector<uint8_t> bytes;
    if (!read_file(manifest_path, bytes)) return segs;
    Reader r(bytes.data(), bytes.size());
    uint32_t n = r.u32();
    if (!r.ok) return segs;
    for (uint32_t i = 0; i < n; i++) {
        std::string s = r.str();
        if (!r.ok) break;  // the reference would yield empty names; we stop at the truncation
        segs.push_back(std::move(s));
    }
    return segs;
}

bool save_manifest(const std::string& manifest_path, const std::vector<std::string>& segs) {
    Writer w;
    w.u32((uint32_t)segs.size());
    for (auto& s : segs) w.str(s);
    return write_file(manifest_path, w.buf.data(), w.buf.size());
}

std::vector<std::string> discover_segments(const std::string& index_dir) {
    std::vector<std::string> names = load_manifest(index_dir + "/manifest.bin");
    if (!names.empty()) return names;
    std::string segroot = index_dir + "/segments";
    DIR* d = opendir(segroot.c_str());
    if (!d) return names;
    while (struct dirent* e = readdir(d)) {
        std::string nm = e->d_name;
        if (nm.rfind("seg_", 0) != 0) continue;
        if (!is_dir(segroot + "/" + nm)) continue;
        names.push_back(nm);
    }
    closedir(d);
    std::sort(names.begin(), names.end());
    return names;
}

namespace {

// One lexicon file: u32 count, then count x {str term, u32 termId, u32 df, u64 offset, u32 count}
// (src/api_segment.cpp:50-61 / :88-99).  `base` is the posting index where this file's
// inverted file starts in the concatenated array; inv_bytes its size for validation.
bool parse_lexicon(const std::vector<uint8_t>& bytes, uint32_t barrel, uint64_t base, uint64_t inv_bytes,
                   std::vector<LexRow>& rows, std::vector<std::string>& terms, const std::string& what) {
    Reader r(bytes.data(), bytes.size());
    uint32_t tcount = r.u32();
    if (!r.ok) { set_error("truncated lexicon header: " + what); return false; }
    // an entry is at least 24 bytes (empty term): a count the file cannot hold is a corrupt header, not a reason to
    // reserve gigabytes
    if ((uint64_t)tcount * 24u > (uint64_t)(r.end - r.p)) { set_error("lexicon entry count exceeds the file size: " + what); return false; }
    rows.reserve(rows.size() + tcount);
    terms.reserve(terms.size() + tcount);
    for (uint32_t i = 0; i < tcount; i++) {
        std::string term = r.str();
        LexRow e;
        e.termId = r.u32();
        e.df = r.u32();
        uint64_t off = r.u64();
        e.count = r.u32();
        e.barrel = barrel;
        if (!r.ok) { set_error("truncated lexicon entry: " + what); return false; }
        if (off % 8 != 0 || off + (uint64_t)e.count * 8 > inv_bytes) {
            set_error("lexicon entry points outside its inverted file (or is not 8-byte aligned): " + what);
            return false;
        }
        e.begin = base + off / 8;
        rows.push_back(e);
        terms.push_back(std::move(term));
    }
    return true;
}

bool slurp_into(const std::string& path, uint8_t* dst, uint64_t at, uint64_t nbytes) {
    FILE* f = std::fopen(path.c_str(), "rb");
    if (!f) return false;
    size_t got = nbytes ? std::fread(dst + at * 8, 1, nbytes, f) : 0;
    std::fclose(f);
    return got == nbytes;
}

uint64_t file_size(const std::string& path, bool& ok) {
    FILE* f = std::fopen(path.c_str(), "rb");
    if (!f) { ok = false; return 0; }
    std::fseek(f, 0, SEEK_END);
    long n = std::ftell(f);
    std::fclose(f);
    ok = n >= 0;
    return ok ? (uint64_t)n : 0;
}

}  // namespace

bool load_segment(const std::string& segdir, HostSegment& s, int nthreads, PostingSink* sink) {
    s = HostSegment{};
    s.dir = segdir;
    std::vector<uint8_t> bytes;

    // stats.bin: u32 N, f32 avgdl  (src/api_segment.cpp:110-115)
    if (!read_file(segdir + "/stats.bin", bytes)) { set_error("cannot open " + segdir + "/stats.bin"); return false; }
    {
        Reader r(bytes.data(), bytes.size());
        s.N = r.u32();
        s.avgdl = r.f32();
        if (!r.ok) { set_error("truncated stats.bin in " + segdir); return false; }
    }

    // docs.bin: u32 n, n x {str cord_uid, str title, str relpath, u32 doc_len}  (:118-131)
    if (!read_file(segdir + "/docs.bin", bytes)) { set_error("cannot open " + segdir + "/docs.bin"); return false; }
    {
        Reader r(bytes.data(), bytes.size());
        uint32_t n = r.u32();
        if (!r.ok) { set_error("truncated docs.bin in " + segdir); return false; }
        // a doc record is at least 16 bytes (three empty strings + doc_len)
        if ((uint64_t)n * 16u > (uint64_t)(r.end - r.p)) { set_error("docs.bin document count exceeds the file size in " + segdir); return false; }
        s.doc_len.resize(n);
        s.uid_off.resize((size_t)n + 1);
        s.uid_chars.reserve((size_t)n * 12);
        s.uid_off[0] = 0;
        for (uint32_t i = 0; i < n; i++) {
            uint32_t len = r.u32();
            if (!r.ok || !r.need(len)) { set_error("truncated docs.bin in " + segdir); return false; }
            s.uid_chars.insert(s.uid_chars.end(), (const char*)r.p, (const char*)r.p + len);
            r.p += len;
            s.uid_off[i + 1] = s.uid_chars.size();
            r.skip_str();  // title
            r.skip_str();  // json_relpath
            s.doc_len[i] = r.u32();
            if (!r.ok) { set_error("truncated docs.bin in " + segdir); return false; }
        }
    }

    std::vector<std::string> terms;
    const bool barrels = file_exists(segdir + "/barrels.bin") && file_exists(segdir + "/inverted_b000.bin") &&
                         file_exists(segdir + "/lexicon_b000.bin");  // include/barrels.hpp:67-71
    if (barrels) {
        s.use_barrels = true;
        if (!read_file(segdir + "/barrels.bin", bytes)) { set_error("cannot open barrels.bin in " + segdir); return false; }
        Reader r(bytes.data(), bytes.size());
        s.barrel_count = r.u32();
        s.terms_per_barrel = r.u32();
        if (!r.ok || s.barrel_count == 0 || s.barrel_count > 65536) { set_error("bad barrels.bin in " + segdir); return false; }
        const uint32_t B = s.barrel_count;
        // every inverted and lexicon barrel must open (src/api_segment.cpp:75-86)
        std::vector<uint64_t> inv_bytes(B), base(B + 1, 0);
        for (uint32_t b = 0; b < B; b++) {
            bool ok = true;
            inv_bytes[b] = file_size(segdir + "/inverted_b" + barrel_suffix(b) + ".bin", ok);
            if (!ok) { set_error("cannot open inverted barrel " + std::to_string(b) + " in " + segdir); return false; }
            if (inv_bytes[b] % 8 != 0) { set_error("inverted barrel size is not a multiple of 8 in " + segdir); return false; }
            base[b + 1] = base[b] + inv_bytes[b] / 8;
        }
        uint8_t* dst = nullptr;
        if (sink) {
            dst = sink->begin(base[B]);
            if (!dst) return false;  // the sink set the error text
        } else {
            s.postings.resize(base[B]);
            dst = (uint8_t*)s.postings.data();
        }
        // largest barrels first: barrel 0 holds the most frequent terms (about half of all postings)
        std::vector<uint32_t> by_size(B);
        for (uint32_t b = 0; b < B; b++) by_size[b] = b;
        std::stable_sort(by_size.begin(), by_size.end(), [&](uint32_t x, uint32_t y) { return inv_bytes[x] > inv_bytes[y]; });
        std::atomic<uint32_t> next{0};
        std::atomic<bool> fail{false};
        auto work = [&]() {
            for (;;) {
                uint32_t i = next.fetch_add(1);
                if (i >= B) break;
                const uint32_t b = by_size[i];
                if (!slurp_into(segdir + "/inverted_b" + barrel_suffix(b) + ".bin", dst, base[b], inv_bytes[b])) fail = true;
                else if (sink) sink->filled(base[b], inv_bytes[b] / 8);
            }
        };
        int nt = std::max(1, std::min(nthreads, (int)B));
        std::vector<std::thread> th;
        for (int t = 1; t < nt; t++) th.emplace_back(work);
        work();
        for (auto& t : th) t.join();
        if (fail) { set_error("short read of an inverted barrel in " + segdir); return false; }
        for (uint32_t b = 0; b < B; b++) {
            std::string lp = segdir + "/lexicon_b" + barrel_suffix(b) + ".bin";
            if (!read_file(lp, bytes)) { set_error("cannot open " + lp); return false; }
            if (!parse_lexicon(bytes, b, base[b], inv_bytes[b], s.rows, terms, lp)) return false;
        }
    } else {
        // legacy: lexicon.bin + inverted.bin (src/api_segment.cpp:45-67)
        s.use_barrels = false;
        std::string lp = segdir + "/lexicon.bin";
        if (!read_file(lp, bytes)) { set_error("cannot open " + lp); return false; }
        bool ok = true;
        uint64_t nb = file_size(segdir + "/inverted.bin", ok);
        if (!ok) { set_error("cannot open " + segdir + "/inverted.bin"); return false; }
        if (nb % 8 != 0) { set_error("inverted.bin size is not a multiple of 8 in " + segdir); return false; }
        uint8_t* dst = nullptr;
        if (sink) {
            dst = sink->begin(nb / 8);
            if (!dst) return false;
        } else {
            s.postings.resize(nb / 8);
            dst = (uint8_t*)s.postings.data();
        }
        if (!slurp_into(segdir + "/inverted.bin", dst, 0, nb)) { set_error("short read of inverted.bin in " + segdir); return false; }
        if (sink) sink->filled(0, nb / 8);
        if (!parse_lexicon(bytes, 0, 0, nb, s.rows, terms, lp)) return false;
    }

    s.lex.reserve(terms.size() * 2 + 1);
    for (uint32_t i = 0; i < (uint32_t)terms.size(); i++) {
        s.rows[i].idf = bm25_idf(s.N, s.rows[i].df);
        s.lex.emplace(std::move(terms[i]), i);  // emplace: first entry for a term wins
    }
    return true;
}

void TermDict::build(const std::vector<std::unique_ptr<HostSegment>>& segs) {
    owned.clear();
    for (size_t i = 0; i < segs.size(); i++)
        if (segs[i]) owned.push_back((uint32_t)i);
    size_t upper = 0;
    for (uint32_t i : owned) upper += segs[i]->lex.size();
    size_t cap = 16;
    while (cap < upper * 2 + 1) cap <<= 1;  // sized for the worst case (disjoint vocabularies): no rehash
    slots.assign(cap, Slot{});
    mask = cap - 1;
    keys.clear();
    nterms = 0;
    // pass 1: intern every term, remember per segment the gid of each row
    std::vector<std::vector<uint32_t>> gid_of(owned.size());
    for (size_t j = 0; j < owned.size(); j++) {
        const HostSegment& sg = *segs[owned[j]];
        gid_of[j].assign(sg.rows.size(), kAbsent);
        for (auto& kv : sg.lex) {
            const std::string& term = kv.first;
            const uint64_t h = term_hash(term.data(), term.size());
            size_t i = (size_t)(h & mask);
            for (;; i = (i + 1) & mask) {
                Slot& sl = slots[i];
                if (sl.gid == kAbsent) {
                    sl.h = h;
                    sl.key_len = (uint32_t)term.size();
                    sl.gid = nterms++;
                    if (term.size() <= sizeof(sl.inl)) {
                        std::memcpy(sl.inl, term.data(), term.size());
                    } else {
                        sl.key_off = (uint32_t)keys.size();
                        keys.insert(keys.end(), term.begin(), term.end());
                    }
                    break;
                }
                if (sl.h == h && key_equal(sl, keys.data(), term.data(), term.size())) break;
            }
            gid_of[j][kv.second] = slots[i].gid;
        }
    }
    // pass 2: the dense (term, segment) table
    table.assign((size_t)nterms * owned.size(), Entry{kAbsent, 0.0f, 0u});
    for (size_t j = 0; j < owned.size(); j++) {
        const HostSegment& sg = *segs[owned[j]];
        for (size_t r = 0; r < sg.rows.size(); r++) {
            const uint32_t g = gid_of[j][r];
            if (g == kAbsent || sg.rows[r].df == 0) continue;  // shadowed duplicate row, or df == 0
            table[(size_t)g * owned.size() + j] = Entry{(uint32_t)r, sg.rows[r].idf, sg.rows[r].count};
        }
    }
}

}  // namespace nsb
