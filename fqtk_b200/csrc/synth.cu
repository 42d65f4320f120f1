// synth.cu — deterministic synthetic workload for the demux matcher.
//
// Counter-based: read i is a pure function of (seed, i, panel), produced by the SAME code on host and device
// (synth_read below is __host__ __device__), so multi-GB batches are generated straight into HBM while tests and
// the CPU baseline regenerate any sub-range on the host.  This is a workload generator, not a matcher: nothing
// here decides assignments.
#include <cstring>
#include <string>
#include <vector>

#include "common.cuh"
#include "kernels.h"

namespace fq {

FQ_HD uint64_t mix64(uint64_t z) {  // splitmix64 finaliser
    z ^= z >> 30;
    z *= 0xBF58476D1CE4E5B9ull;
    z ^= z >> 27;
    z *= 0x94D049BB133111EBull;
    z ^= z >> 31;
    return z;
}

struct Rng {
    uint64_t s;
    FQ_HD uint64_t next() {
        s += 0x9E3779B97F4A7C15ull;
        return mix64(s);
    }
};

FQ_HD uint32_t bounded(uint64_t r, uint32_t n) { return (uint32_t)(((r >> 32) * (uint64_t)n) >> 32); }

FQ_HD uint32_t popc4(uint32_t m) { return (m & 1u) + ((m >> 1) & 1u) + ((m >> 2) & 1u) + ((m >> 3) & 1u); }

// index (0..3 = A,C,G,T) of the t-th set bit of a 4-bit base set
FQ_HD uint32_t nth_base(uint32_t m, uint32_t t) {
    for (uint32_t b = 0; b < 4u; b++) {
        if ((m >> b) & 1u) {
            if (t == 0u) return b;
            t--;
        }
    }
    return 0u;
}

// Writes the L ASCII bases of read `i` to out[0..L).
FQ_HD void synth_read(const uint8_t* panel, uint32_t S, uint32_t L, uint64_t seed, uint64_t i, uint8_t* out) {
    const char BASES[4] = {'A', 'C', 'G', 'T'};
    Rng g{mix64(seed ^ (i * 0xD1B54A32D192ED03ull))};
    const uint32_t kind = bounded(g.next(), 100u);
    if (kind >= 98u) {  // 2 %: uniform random L-mer (index hopping / PhiX stand-in)
        uint64_t r = 0;
        for (uint32_t k = 0; k < L; k++) {
            if ((k & 31u) == 0u) r = g.next();
            out[k] = (uint8_t)BASES[r & 3u];
            r >>= 2;
        }
        return;
    }
    const uint32_t j = bounded(g.next(), S);
    for (uint32_t k = 0; k < L; k++) {
        uint32_t m = encode_byte(panel[(size_t)j * L + k]);
        if (m == 0u) m = 15u;
        const uint64_t r = g.next();
        const uint32_t n = popc4(m);
        uint32_t b = nth_base(m, (uint32_t)(((r >> 40) * (uint64_t)n) >> 24));  // resolve degenerate positions
        if ((r & 0xFFFFu) < 328u) b = (b + 1u + (uint32_t)((r >> 16) & 0xFFu) % 3u) & 3u;  // 0.5 % substitution
        out[k] = (uint8_t)BASES[b];
        if (((r >> 24) & 0xFFFFu) < 131u) out[k] = 'N';  // 0.2 % no-call
    }
    if (kind >= 90u) {  // 8 %: near-miss, 2-3 forced substitutions at distinct positions
        uint32_t nsub = 2u + (uint32_t)(g.next() & 1u);
        if (nsub > L) nsub = L;
        uint32_t pos[3] = {0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu};
        for (uint32_t s = 0; s < nsub; s++) {
            uint32_t pk;
            uint64_t r;
            do {
                r = g.next();
                pk = bounded(r, L);
            } while (pk == pos[0] || pk == pos[1] || pk == pos[2]);
            pos[s] = pk;
            const uint8_t c = out[pk];
            const uint32_t cur = c == 'C' ? 1u : c == 'G' ? 2u : c == 'T' ? 3u : 0u;
            out[pk] = (uint8_t)BASES[(cur + 1u + (uint32_t)(r & 0xFFFFu) % 3u) & 3u];
        }
    }
}

__global__ void __launch_bounds__(256) k_synth(const uint8_t* __restrict__ panel, uint32_t S, uint32_t L,
                                               uint64_t seed, uint64_t first, uint64_t n,
                                               uint8_t* __restrict__ ascii, uint32_t* __restrict__ packed) {
    const uint32_t W = words_for_len(L);
    const uint64_t total = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; t < n; t += total) {
        uint8_t buf[256];
        synth_read(panel, S, L, seed, first + t, buf);
        if (ascii)
            for (uint32_t k = 0; k < L; k++) ascii[t * L + k] = buf[k];
        if (packed) {
            for (uint32_t wi = 0; wi < W; wi++) {
                uint32_t acc = 0u;
                for (uint32_t b = 0; b < 8u; b++) {
                    const uint32_t k = wi * 8u + b;
                    if (k < L) acc |= encode_byte(buf[k]) << (4u * b);
                }
                packed[t * W + wi] = acc;
            }
        }
    }
}

cudaError_t synth_reads_device(const uint8_t* d_panel, uint32_t S, uint32_t L, uint64_t seed, uint64_t first,
                               uint64_t n, uint8_t* d_ascii, uint32_t* d_packed, int sm_count, cudaStream_t stream) {
    if (n == 0) return cudaSuccess;
    uint64_t want = (n + 255) / 256;
    const uint64_t cap = (uint64_t)sm_count * 8;
    const int grid = (int)(want < cap ? want : cap);
    k_synth<<<grid, 256, 0, stream>>>(d_panel, S, L, seed, first, n, d_ascii, d_packed);
    count_launch();
    return cudaGetLastError();
}

void synth_reads_host(const uint8_t* panel, uint32_t S, uint32_t L, uint64_t seed, uint64_t first, uint64_t n,
                      uint8_t* out) {
    for (uint64_t t = 0; t < n; t++) synth_read(panel, S, L, seed, first + t, out + t * L);
}

// Panel: greedy rejection on pairwise Hamming distance, then optional degenerate rewriting (cfg 5).
int synth_panel_host(uint64_t seed, uint32_t S, uint32_t L, uint32_t min_distance, uint32_t n_degenerate,
                     uint8_t* out) {
    static const char BASES[4] = {'A', 'C', 'G', 'T'};
    static const char CODES[] = "RYSWKMBDHVN";
    std::vector<uint8_t> acc;  // accepted, 1 byte per base (0..3)
    acc.reserve((size_t)S * L);
    std::vector<uint8_t> cand(L);
    uint64_t counter = 0, rejected = 0;
    uint32_t have = 0;
    while (have < S) {
        Rng g{mix64(seed + 0x5851F42D4C957F2Dull * (++counter))};
        uint64_t r = 0;
        for (uint32_t k = 0; k < L; k++) {
            if ((k & 31u) == 0u) r = g.next();
            cand[k] = (uint8_t)(r & 3u);
            r >>= 2;
        }
        bool ok = true;
        for (uint32_t j = 0; j < have && ok; j++) {
            uint32_t d = 0;
            const uint8_t* o = acc.data() + (size_t)j * L;
            for (uint32_t k = 0; k < L && d < min_distance; k++) d += (o[k] != cand[k]);
            ok = d >= min_distance;
        }
        if (!ok) {
            if (++rejected > 50000000ull) return -1;  // panel infeasible for this (S, L, min_distance)
            continue;
        }
        acc.insert(acc.end(), cand.begin(), cand.end());
        have++;
    }
    for (size_t t = 0; t < (size_t)S * L; t++) out[t] = (uint8_t)BASES[acc[t]];
    if (n_degenerate > 0) {
        if (n_degenerate > L) n_degenerate = L;
        for (uint32_t j = 0; j < S; j++) {
            uint8_t* bc = out + (size_t)j * L;
            std::vector<uint8_t> orig(bc, bc + L);
            for (uint64_t attempt = 0;; attempt++) {
                std::memcpy(bc, orig.data(), L);
                Rng g{mix64(seed ^ (0xA24BAED4963EE407ull * (j + 1)) ^ (attempt << 40))};
                std::vector<uint32_t> used;
                while (used.size() < n_degenerate) {
                    const uint64_t r = g.next();
                    const uint32_t pk = bounded(r, L);
                    bool dup = false;
                    for (uint32_t u : used) dup = dup || (u == pk);
                    if (dup) continue;
                    const uint32_t have_mask = encode_byte(orig[pk]);
                    // uniformly among the degenerate codes that still admit the original base
                    uint32_t n_ok = 0;
                    for (int c = 0; c < 11; c++) n_ok += (encode_byte((uint8_t)CODES[c]) & have_mask) != 0u;
                    uint32_t pick = (uint32_t)((r & 0xFFFFu) % n_ok);
                    for (int c = 0; c < 11; c++) {
                        if ((encode_byte((uint8_t)CODES[c]) & have_mask) != 0u) {
                            if (pick == 0u) {
                                bc[pk] = (uint8_t)CODES[c];
                                break;
                            }
                            pick--;
                        }
                    }
                    used.push_back(pk);
                }
                bool unique = true;  // samples.rs:112-115: barcodes must be unique as strings
                for (uint32_t o = 0; o < j && unique; o++) unique = std::memcmp(out + (size_t)o * L, bc, L) != 0;
                if (unique) break;
                if (attempt > 1000) return -1;
            }
        }
    }
    return 0;
}

}  // namespace fq
