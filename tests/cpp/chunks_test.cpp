// CPU unit test of the batch chunking rule (superodom_b200/csrc/so_chunks.h): prints the bounds for the sizes on the command
// line and checks the invariants for every batch size up to 2048 in every mode.
#include <cstdio>
#include <cstdlib>

#include "so_chunks.h"

static int check(size_t n, bool host, bool prof, int ov) {
    const std::vector<uint32_t> b = so::chunk_bounds(n, host, prof, ov);
    if (b.size() < 2 || b.front() != 0 || b.back() != n) return 1;                       // covers [0, n) ...
    if (b.size() - 1 > so::kMaxChunks) return 2;                                         // ... in at most 16 chunks
    for (size_t k = 1; k < b.size(); ++k) if (b[k] < b[k - 1]) return 3;                 // ... in order (equal = empty chunk, skipped by the caller)
    if (prof && b.size() != 2) return 4;
    if (!prof && ov == 0) {
        for (size_t k = 1; k < b.size(); ++k) if (b[k] == b[k - 1] && n) return 5;       // the built-in rules make no empty chunks
        if (host && n >= 32) {
            if (b[1] != 8) return 6;                                                     // only an 8-scan upload is exposed
            for (size_t k = 1; k + 1 < b.size(); ++k) if (b[k + 1] - b[k] < 8) return 7; // no sliver: every later chunk holds >= 8 scans
            for (size_t k = 2; k + 1 < b.size(); ++k) if (b[k + 1] - b[k] < b[k] - b[k - 1] && k + 2 < b.size()) return 8;   // chunks grow (the last one may be a remainder)
        } else if (b.size() > 3) return 9;
    }
    return 0;
}

int main(int argc, char** argv) {
    for (int i = 1; i < argc; ++i) {
        const size_t n = size_t(std::atoi(argv[i]));
        const std::vector<uint32_t> b = so::chunk_bounds(n, true, false, 0);
        std::printf("%zu:", n);
        for (uint32_t v : b) std::printf(" %u", v);
        std::printf("\n");
    }
    for (size_t n = 1; n <= 2048; ++n)
        for (int host = 0; host < 2; ++host)
            for (int prof = 0; prof < 2; ++prof)
                for (int ov : {0, 1, 2, 4, 7, 16, 40}) {
                    const int rc = check(n, host != 0, prof != 0, ov);
                    if (rc) { std::printf("FAIL n=%zu host=%d prof=%d override=%d rule %d\n", n, host, prof, ov, rc); return 1; }
                }
    std::printf("ok\n");
    return 0;
}
