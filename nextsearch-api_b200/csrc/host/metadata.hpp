// Result decoration from INDEX_DIR/metadata.csv.
//
// The reference keeps, per cord_uid, the byte position of the FIRST csv line that names it
// (load_metadata_uid_meta, src/api_metadata.cpp:112-186) and, for every hit of every search, reopens the
// file, re-reads that line and the header line and re-derives the column positions
// (fetch_metadata, :189-249; called from src/api_engine.cpp:516-532).  Here the file is read once at
// reload, the header is parsed once, the uid -> line table is a flat hash, and decorating a hit is a parse
// of one in-memory line.  Same observable fields:
//   * a "line" is what std::getline yields: bytes up to '\n' (a '\r' before it stays in the last column);
//   * columns: split at ',' outside double quotes; every '"' toggles the quoted state and is dropped
//     (no "" escape) — csv_row, :12-42;
//   * when a header name occurs twice the LAST position wins (:137-139, :222-228);
//   * rows with too few columns or an empty uid are skipped; the first row of a uid wins (:166-172);
//   * author = surname of the first author + " et al." (first_author_et_al, :58-110).
#pragma once
#include <cstdint>
#include <cstring>
#include <string>
#include <vector>

#include "common.hpp"
#include "segment_io.hpp"  // term_hash

namespace nsb {

struct MetaFields {
    std::string title, url, publish_time, author;  // raw column values (url before the ';' cut of api_engine.cpp:524-526)
};

class MetaIndex {
  public:
    // false when the file is absent or has no cord_uid column (the reference only logs a warning)
    bool load(const std::string& csv_path) {
        bytes_.clear();
        slots_.clear();
        nrows_ = 0;
        uid_i_ = url_i_ = pub_i_ = auth_i_ = title_i_ = -1;
        if (!read_file(csv_path, bytes_) || bytes_.empty()) return false;
        const char* p = (const char*)bytes_.data();
        const size_t n = bytes_.size();
        size_t eol = line_end(0);
        std::vector<std::string> cols;
        split(p, eol, cols);
        for (int i = 0; i < (int)cols.size(); i++) {
            if (cols[i] == "cord_uid") uid_i_ = i;
            if (cols[i] == "url") url_i_ = i;
            if (cols[i] == "publish_time") pub_i_ = i;
            if (cols[i] == "authors") auth_i_ = i;
            if (cols[i] == "title") title_i_ = i;
        }
        if (uid_i_ < 0) return false;
        // pass 1: count lines to size the table; pass 2: insert
        size_t lines = 0;
        for (size_t at = eol + 1; at < n; at = line_end(at) + 1) lines++;
        size_t cap = 16;
        while (cap < lines * 2 + 1) cap <<= 1;
        slots_.assign(cap, Slot{});
        mask_ = cap - 1;
        std::string uid;
        for (size_t at = eol + 1; at < n;) {
            const size_t e = line_end(at);
            if (column(p + at, e - at, uid_i_, uid) && !uid.empty()) insert(uid, at, e - at);
            at = e + 1;
        }
        return true;
    }

    bool enabled() const { return !slots_.empty(); }
    size_t rows() const { return nrows_; }

    bool lookup(const std::string& uid, MetaFields& out) const {
        if (slots_.empty()) return false;
        const uint64_t h = term_hash(uid.data(), uid.size());
        const char* base = (const char*)bytes_.data();
        for (size_t i = (size_t)(h & mask_);; i = (i + 1) & mask_) {
            const Slot& s = slots_[i];
            if (s.len == kEmpty) return false;
            if (s.h != h) continue;
            std::string u;
            if (!column(base + s.off, s.len, uid_i_, u) || u != uid) continue;
            std::string authors;
            out = MetaFields{};
            if (url_i_ >= 0) column(base + s.off, s.len, url_i_, out.url);
            if (pub_i_ >= 0) column(base + s.off, s.len, pub_i_, out.publish_time);
            if (title_i_ >= 0) column(base + s.off, s.len, title_i_, out.title);
            if (auth_i_ >= 0 && column(base + s.off, s.len, auth_i_, authors)) out.author = first_author(authors);
            return true;
        }
    }

    // surname of the first author + " et al." ("" when there is none)
    static std::string first_author(const std::string& raw) {
        std::string s = trimmed(raw);
        if (s.empty()) return s;
        const size_t semi = s.find(';');
        std::string first = trimmed(semi == std::string::npos ? s : s.substr(0, semi));
        while (!first.empty() && (first.back() == ',' || is_space((unsigned char)first.back()))) first.pop_back();
        first = trimmed(first);
        if (first.empty()) return first;
        if (first.front() == '(') {  // romanised name in parentheses
            const size_t close = first.find(')');
            if (close != std::string::npos && close > 1) {
                std::string inside = trimmed(first.substr(1, close - 1));
                if (!inside.empty()) first = inside;
            }
        }
        std::string surname;
        const size_t comma = first.find(',');
        if (comma != std::string::npos) {
            surname = trimmed(first.substr(0, comma));
        } else {
            const size_t sp = first.find_last_of(" \t");
            surname = sp == std::string::npos ? first : trimmed(first.substr(sp + 1));
        }
        surname = trimmed(surname);
        return surname.empty() ? surname : surname + " et al.";
    }

  private:
    static constexpr uint32_t kEmpty = 0xFFFFFFFFu;
    struct Slot {
        uint64_t h = 0;
        uint64_t off = 0;
        uint32_t len = kEmpty;
        uint32_t pad = 0;
    };
    static bool is_space(unsigned char c) { return c == ' ' || (c >= '\t' && c <= '\r'); }  // std::isspace, "C" locale
    static std::string trimmed(const std::string& s) {
        size_t a = 0, b = s.size();
        while (a < b && is_space((unsigned char)s[a])) a++;
        while (b > a && is_space((unsigned char)s[b - 1])) b--;
        return s.substr(a, b - a);
    }
    size_t line_end(size_t at) const {
        const void* nl = std::memchr(bytes_.data() + at, '\n', bytes_.size() - at);
        return nl ? (size_t)((const uint8_t*)nl - bytes_.data()) : bytes_.size();
    }
    static void split(const char* p, size_t n, std::vector<std::string>& out) {
        out.clear();
        std::string cur;
        bool inq = false;
        for (size_t i = 0; i < n; i++) {
            const char c = p[i];
            if (c == '"') inq = !inq;
            else if (!inq && c == ',') { out.push_back(cur); cur.clear(); }
            else cur.push_back(c);
        }
        out.push_back(cur);
    }
    // value of column `want` of one line; false when the line has fewer columns
    static bool column(const char* p, size_t n, int want, std::string& out) {
        out.clear();
        bool inq = false;
        int col = 0;
        for (size_t i = 0; i < n; i++) {
            const char c = p[i];
            if (c == '"') inq = !inq;
            else if (!inq && c == ',') {
                if (col == want) return true;
                col++;
            } else if (col == want) out.push_back(c);
        }
        return col == want;
    }
    void insert(const std::string& uid, size_t off, size_t len) {
        const uint64_t h = term_hash(uid.data(), uid.size());
        const char* base = (const char*)bytes_.data();
        std::string other;
        for (size_t i = (size_t)(h & mask_);; i = (i + 1) & mask_) {
            Slot& s = slots_[i];
            if (s.len == kEmpty) {
                s.h = h;
                s.off = off;
                s.len = (uint32_t)len;
                nrows_++;
                return;
            }
            if (s.h == h && column(base + s.off, s.len, uid_i_, other) && other == uid) return;  // first row wins
        }
    }
    std::vector<uint8_t> bytes_;
    std::vector<Slot> slots_;
    uint64_t mask_ = 0;
    size_t nrows_ = 0;
    int uid_i_ = -1, url_i_ = -1, pub_i_ = -1, auth_i_ = -1, title_i_ = -1;
};

}  // namespace nsb
