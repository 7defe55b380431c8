#include "common.hpp"

#include <sys/stat.h>
#include <sys/types.h>

#include <cerrno>

namespace nsb {

static thread_local std::string g_err;

void set_error(const std::string& msg) { g_err = msg; }
const char* last_error() { return g_err.c_str(); }

bool read_file(const std::string& path, std::vector<uint8_t>& out) {
    FILE* f = std::fopen(path.c_str(), "rb");
    if (!f) return false;
    struct stat st;
    if (fstat(fileno(f), &st) != 0) { std::fclose(f); return false; }
    out.resize((size_t)st.st_size);
    size_t got = out.empty() ? 0 : std::fread(out.data(), 1, out.size(), f);
    std::fclose(f);
    return got == out.size();
}

bool write_file(const std::string& path, const void* data, size_t n) {
    FILE* f = std::fopen(path.c_str(), "wb");
    if (!f) return false;
    size_t put = n ? std::fwrite(data, 1, n, f) : 0;
    bool ok = (put == n) && (std::fclose(f) == 0);
    return ok;
}

bool file_exists(const std::string& path) {
    struct stat st;
    return stat(path.c_str(), &st) == 0;
}

bool is_dir(const std::string& path) {
    struct stat st;
    return stat(path.c_str(), &st) == 0 && S_ISDIR(st.st_mode);
}

bool make_dirs(const std::string& path) {
    if (path.empty()) return false;
    std::string cur;
    for (size_t i = 0; i <= path.size(); i++) {
        if (i == path.size() || path[i] == '/') {
            if (!cur.empty() && !is_dir(cur)) {
                if (mkdir(cur.c_str(), 0777) != 0 && errno != EEXIST) return false;
            }
        }
        if (i < path.size()) cur.push_back(path[i]);
    }
    return is_dir(path);
}

}  // namespace nsb
