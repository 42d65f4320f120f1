// emit_kernels.cu — the routed records written out on the device: SampleWriters::write (src/bin/commands/demux.rs:396-415)
// and ReadSet::write_header_internal (:171-267) for a whole batch (SURVEY.md 8f "next" #3 / #4).
//
// After matching and routing, the reference walks the read sets one by one and hands every output segment to its
// sample's writer: header (rewritten: read number, UMIs appended to the name, sample barcode appended to the comment's
// index field), '\n', bases, "\n+\n", qualities, '\n'.  Here the batch's FASTQ chunks are already in device memory
// (ingest_kernels.cu found the records, route_kernels.cu put the read indices in per-sample, input-ordered runs), so the
// same bytes are produced there: one text buffer per output stream (R1, R2, I1, U1, ...), every sample's records one
// contiguous run in it, ready for the BGZF kernel — the host only learns where each (stream, sample) run starts.
//   k_emit_lengths   per (stream, routed position): bytes of the record; header rule violations -> first offender
//   k_emit_tile_sums / k_emit_scan_tiles / k_emit_apply   exclusive prefix sums of the lengths = where every record starts
//   k_emit_write     the records
//   k_emit_file_offsets   start of every (stream, bucket) run
#include <cuda_runtime.h>

#include <algorithm>
#include <cstdint>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/fqtk_b200.h"
#include "kernels.h"

namespace fq {
void set_last_error(const std::string& msg);
}

namespace {

constexpr uint32_t EM_MAX_SRC = 8, EM_MAX_SEG = 32, EM_MAX_STREAMS = 16;
constexpr uint32_t EM_REST = 0xFFFFFFFFu;
constexpr int EM_THREADS = 256;
constexpr uint32_t EM_TILE = 1024;  // lengths per CTA in the scan kernels

struct EmitSeg {
    uint32_t source, kind, offset, length;
};
struct EmitPlan {
    const uint8_t* chunk[EM_MAX_SRC];
    const unsigned long long* head_off[EM_MAX_SRC];
    const unsigned long long* seq_off[EM_MAX_SRC];
    const uint32_t* seq_len[EM_MAX_SRC];
    EmitSeg seg[EM_MAX_SEG];
    uint32_t n_sources, n_segs, n_streams;
    uint32_t stream_seg[EM_MAX_STREAMS];      // the segment a stream writes
    uint32_t stream_readnum[EM_MAX_STREAMS];  // 1-based index of that segment among the segments of its kind
};

// one segment of read i: pointer to its bases and its length (a REST segment runs to the end of the line)
struct SegView {
    const uint8_t* bases;
    uint32_t len;
    uint64_t seq_off;  // of the source's record (for the quality line)
    uint32_t seq_len;
    uint32_t offset;
};
__device__ __forceinline__ SegView seg_view(const EmitPlan& p, uint32_t k, uint64_t i) {
    const EmitSeg& s = p.seg[k];
    SegView v;
    v.seq_off = p.seq_off[s.source][i];
    v.seq_len = p.seq_len[s.source][i];
    v.offset = s.offset;
    v.bases = p.chunk[s.source] + v.seq_off + s.offset;
    v.len = s.length == EM_REST ? (v.seq_len > s.offset ? v.seq_len - s.offset : 0u) : s.length;
    return v;
}
// first byte of the quality line of a record whose sequence line is [seq_off, seq_off + seq_len)
__device__ __forceinline__ const uint8_t* qual_line(const uint8_t* chunk, uint64_t seq_off, uint32_t seq_len) {
    const uint8_t* q = chunk + seq_off + seq_len;
    while (*q != '\n') q++;  // (a '\r')
    q++;
    while (*q != '\n') q++;  // the '+' line
    return q + 1;
}

// ReadSet::write_header_internal (demux.rs:171-267) as a plan: which pieces of the old header survive
struct HeaderPlan {
    const uint8_t* h;    // header text without '@'
    uint32_t hlen, name_len;
    uint32_t mode;       // 0 no comment: "<n>:N:0:"; 1 comment with < 4 fields: kept, ':' appended if missing;
                         // 2 four fields: read number replaced, index field kept (a lone trailing digit dropped)
    uint32_t c_off, c_len, rem_off, rem_len;
    uint32_t add_colon, add_plus, umi_sep_plus;
    uint32_t err;        // 0 ok, 1 name with more than 8 parts (UMIs present), 2 comment without 4 parts, 3 empty comment
};
__device__ HeaderPlan plan_header(const EmitPlan& p, uint64_t i, bool has_umi) {
    HeaderPlan hp{};
    const uint64_t h0 = p.head_off[0][i], s0 = p.seq_off[0][i];
    hp.h = p.chunk[0] + h0 + 1;
    uint32_t hlen = (uint32_t)(s0 - h0 - 2);
    if (hlen && hp.h[hlen - 1] == '\r') hlen--;
    hp.hlen = hlen;
    uint32_t sp = hlen;
    for (uint32_t k = 0; k < hlen; k++)
        if (hp.h[k] == ' ') { sp = k; break; }
    hp.name_len = sp;
    if (has_umi) {
        uint32_t colons = 0;
        for (uint32_t k = 0; k < sp; k++) colons += hp.h[k] == ':';
        if (colons > 7) hp.err = 1;
        hp.umi_sep_plus = colons == 7;
    }
    if (sp == hlen) {
        hp.mode = 0;
        return hp;
    }
    hp.c_off = sp + 1;
    hp.c_len = hlen - sp - 1;
    if (hp.c_len == 0) {
        if (!hp.err) hp.err = 3;
        return hp;
    }
    uint32_t colons = 0, first_colon = 0;
    for (uint32_t k = 0; k < hp.c_len; k++)
        if (hp.h[hp.c_off + k] == ':') {
            if (!colons) first_colon = k;
            colons++;
        }
    const uint8_t last = hp.h[hlen - 1];
    if (colons < 3) {
        hp.mode = 1;
        hp.add_colon = last != ':';
    } else if (colons != 3) {
        if (!hp.err) hp.err = 2;
    } else {
        hp.mode = 2;
        hp.rem_off = hp.c_off + first_colon + 1;
        const uint32_t rem_end = (last >= '0' && last <= '9') ? hlen - 1 : hlen;
        hp.rem_len = rem_end - hp.rem_off;
        hp.add_plus = hp.rem_len == 0 || hp.h[rem_end - 1] != ':';
    }
    return hp;
}

__device__ __forceinline__ uint32_t digits(uint32_t r) { return r >= 100 ? 3 : (r >= 10 ? 2 : 1); }

// bytes of the record of read i in stream t
__device__ uint32_t record_length(const EmitPlan& p, uint32_t t, uint64_t i, uint32_t& err) {
    uint32_t bc = 0, nb = 0, umi = 0, nm = 0;
    for (uint32_t k = 0; k < p.n_segs; k++) {
        if (p.seg[k].kind == 'B') { bc += seg_view(p, k, i).len; nb++; }
        else if (p.seg[k].kind == 'M') { umi += seg_view(p, k, i).len; nm++; }
    }
    const HeaderPlan hp = plan_header(p, i, nm != 0);
    err = hp.err;
    const uint32_t d = digits(p.stream_readnum[t]);
    uint32_t len = 1 + hp.name_len + (nm ? 1 + umi + (nm - 1) : 0) + 1;
    len += hp.mode == 0 ? d + 5 : (hp.mode == 1 ? hp.c_len + hp.add_colon : d + 1 + hp.rem_len + hp.add_plus);
    len += nb ? bc + (nb - 1) : 0;
    const uint32_t sl = seg_view(p, p.stream_seg[t], i).len;
    return len + 1 + sl + 3 + sl + 1;
}

__global__ void __launch_bounds__(EM_THREADS) k_emit_lengths(const EmitPlan p, const uint32_t* __restrict__ order, uint64_t n,
                                                            uint32_t* __restrict__ lens, unsigned long long* __restrict__ err) {
    const uint32_t t = blockIdx.y;
    for (uint64_t j = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; j < n; j += (uint64_t)gridDim.x * blockDim.x) {
        uint32_t e;
        lens[(size_t)t * n + j] = record_length(p, t, order[j], e);
        if (e && t == 0) atomicMin(err, (unsigned long long)order[j] << 2 | e);
    }
}

__global__ void __launch_bounds__(EM_THREADS) k_emit_tile_sums(const uint32_t* __restrict__ lens, uint64_t n, uint32_t n_tiles,
                                                              unsigned long long* __restrict__ tile_sums) {
    __shared__ unsigned long long s_w[EM_THREADS / 32];
    const uint32_t t = blockIdx.y;
    const uint64_t base = (uint64_t)blockIdx.x * EM_TILE;
    unsigned long long c = 0;
    for (uint32_t k = threadIdx.x; k < EM_TILE; k += EM_THREADS)
        if (base + k < n) c += lens[(size_t)t * n + base + k];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xFFFFFFFFu, c, o);
    if ((threadIdx.x & 31u) == 0) s_w[threadIdx.x >> 5] = c;
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned long long s = 0;
        for (int k = 0; k < EM_THREADS / 32; k++) s += s_w[k];
        tile_sums[(size_t)t * (n_tiles + 1) + blockIdx.x] = s;
    }
}

// one CTA per stream: exclusive scan of its tile sums in place, total behind them
__global__ void __launch_bounds__(1024) k_emit_scan_tiles(unsigned long long* __restrict__ tile_sums, uint32_t n_tiles) {
    __shared__ unsigned long long part[1024];
    unsigned long long* row = tile_sums + (size_t)blockIdx.x * (n_tiles + 1);
    const uint32_t per = (n_tiles + 1023u) / 1024u;
    const uint32_t lo = min(threadIdx.x * per, n_tiles), hi = min(lo + per, n_tiles);
    unsigned long long sum = 0;
    for (uint32_t k = lo; k < hi; k++) sum += row[k];
    part[threadIdx.x] = sum;
    __syncthreads();
    for (uint32_t off = 1; off < 1024u; off <<= 1) {
        const unsigned long long v = threadIdx.x >= off ? part[threadIdx.x - off] : 0ull;
        __syncthreads();
        part[threadIdx.x] += v;
        __syncthreads();
    }
    unsigned long long run = part[threadIdx.x] - sum;
    for (uint32_t k = lo; k < hi; k++) {
        const unsigned long long c = row[k];
        row[k] = run;
        run += c;
    }
    if (threadIdx.x == 1023u) row[n_tiles] = part[1023];
}

// pre[t][j] = bytes of stream t before routed position j; pre[t][n] = total
__global__ void __launch_bounds__(EM_THREADS) k_emit_apply(const uint32_t* __restrict__ lens, uint64_t n, uint32_t n_tiles,
                                                          const unsigned long long* __restrict__ tile_sums,
                                                          unsigned long long* __restrict__ pre) {
    __shared__ unsigned long long s_w[EM_THREADS / 32];
    const uint32_t t = blockIdx.y, lane = threadIdx.x & 31u, w = threadIdx.x >> 5;
    const uint64_t base = (uint64_t)blockIdx.x * EM_TILE + (uint64_t)threadIdx.x * (EM_TILE / EM_THREADS);
    uint32_t v[EM_TILE / EM_THREADS];
    unsigned long long c = 0;
#pragma unroll
    for (uint32_t k = 0; k < EM_TILE / EM_THREADS; k++) {
        v[k] = base + k < n ? lens[(size_t)t * n + base + k] : 0u;
        c += v[k];
    }
    unsigned long long incl = c;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const unsigned long long u = __shfl_up_sync(0xFFFFFFFFu, incl, o);
        if (lane >= (uint32_t)o) incl += u;
    }
    if (lane == 31u) s_w[w] = incl;
    __syncthreads();
    unsigned long long run = tile_sums[(size_t)t * (n_tiles + 1) + blockIdx.x] + incl - c;
    for (uint32_t q = 0; q < w; q++) run += s_w[q];
#pragma unroll
    for (uint32_t k = 0; k < EM_TILE / EM_THREADS; k++) {
        if (base + k < n) pre[(size_t)t * (n + 1) + base + k] = run;
        run += v[k];
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) pre[(size_t)t * (n + 1) + n] = tile_sums[(size_t)t * (n_tiles + 1) + n_tiles];
}

__device__ __forceinline__ uint8_t* put(uint8_t* d, const uint8_t* s, uint32_t n) {
    for (uint32_t k = 0; k < n; k++) d[k] = s[k];
    return d + n;
}
__device__ __forceinline__ uint8_t* put_num(uint8_t* d, uint32_t r) {
    if (r >= 100) *d++ = (uint8_t)('0' + r / 100 % 10);
    if (r >= 10) *d++ = (uint8_t)('0' + r / 10 % 10);
    *d++ = (uint8_t)('0' + r % 10);
    return d;
}

// Thread per record for the header (a few dozen bytes assembled piece by piece), then the WARP copies the bases and the
// qualities of its 32 records one record at a time, 32 consecutive bytes per step: those two lines are ~6/7 of the bytes,
// and written one byte per thread they touch 32 different sectors per store instruction (the first version of this
// kernel: 3.1 ms for 294 MB of text).
__global__ void __launch_bounds__(EM_THREADS) k_emit_write(const EmitPlan p, const uint32_t* __restrict__ order, uint64_t n,
                                                          const unsigned long long* __restrict__ pre,
                                                          const unsigned long long* __restrict__ stream_base,
                                                          uint8_t* __restrict__ text) {
    const uint32_t t = blockIdx.y;
    const uint32_t lane = threadIdx.x & 31u;
    const uint64_t n_round = (n + 31u) & ~(uint64_t)31u;  // whole warps stay in the loop (shuffles below)
    for (uint64_t j = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; j < n_round; j += (uint64_t)gridDim.x * blockDim.x) {
        const uint8_t *src_b = nullptr, *src_q = nullptr;
        uint8_t* dst = nullptr;
        uint32_t len = 0;
        if (j < n) {
            const uint64_t i = order[j];
            uint8_t* d = text + stream_base[t] + pre[(size_t)t * (n + 1) + j];
            uint32_t nm = 0;
            for (uint32_t k = 0; k < p.n_segs; k++) nm += p.seg[k].kind == 'M';
            const HeaderPlan hp = plan_header(p, i, nm != 0);
            *d++ = '@';
            d = put(d, hp.h, hp.name_len);
            if (nm) {
                *d++ = hp.umi_sep_plus ? '+' : ':';
                bool first = true;
                for (uint32_t k = 0; k < p.n_segs; k++)
                    if (p.seg[k].kind == 'M') {
                        if (!first) *d++ = '+';
                        first = false;
                        const SegView v = seg_view(p, k, i);
                        d = put(d, v.bases, v.len);
                    }
            }
            *d++ = ' ';
            const uint32_t r = p.stream_readnum[t];
            if (hp.mode == 0) {
                d = put_num(d, r);
                *d++ = ':'; *d++ = 'N'; *d++ = ':'; *d++ = '0'; *d++ = ':';
            } else if (hp.mode == 1) {
                d = put(d, hp.h + hp.c_off, hp.c_len);
                if (hp.add_colon) *d++ = ':';
            } else {
                d = put_num(d, r);
                *d++ = ':';
                d = put(d, hp.h + hp.rem_off, hp.rem_len);
                if (hp.add_plus) *d++ = '+';
            }
            {
                bool first = true;
                for (uint32_t k = 0; k < p.n_segs; k++)
                    if (p.seg[k].kind == 'B') {
                        if (!first) *d++ = '+';
                        first = false;
                        const SegView v = seg_view(p, k, i);
                        d = put(d, v.bases, v.len);
                    }
            }
            *d++ = '\n';
            const uint32_t sk = p.stream_seg[t];
            const SegView v = seg_view(p, sk, i);
            src_b = v.bases;
            src_q = qual_line(p.chunk[p.seg[sk].source], v.seq_off, v.seq_len) + v.offset;
            dst = d;
            len = v.len;
            d += len;
            *d++ = '\n'; *d++ = '+'; *d++ = '\n';
            d[len] = '\n';
        }
        // bases at dst[0, len), qualities at dst[len + 3, 2 len + 3): record rr of the warp, all lanes
        for (int rr = 0; rr < 32; rr++) {
            const uint32_t ln = __shfl_sync(0xFFFFFFFFu, len, rr);
            if (ln == 0u) continue;  // (warp-uniform)
            const uint8_t* sb = reinterpret_cast<const uint8_t*>(__shfl_sync(0xFFFFFFFFu, reinterpret_cast<unsigned long long>(src_b), rr));
            const uint8_t* sq = reinterpret_cast<const uint8_t*>(__shfl_sync(0xFFFFFFFFu, reinterpret_cast<unsigned long long>(src_q), rr));
            uint8_t* dd = reinterpret_cast<uint8_t*>(__shfl_sync(0xFFFFFFFFu, reinterpret_cast<unsigned long long>(dst), rr));
            for (uint32_t k0 = 0; k0 < ln; k0 += 128u) {  // four steps' loads in flight
                uint8_t b[4], q[4];
#pragma unroll
                for (int u = 0; u < 4; u++) {
                    const uint32_t k = k0 + 32u * u + lane;
                    if (k < ln) {
                        b[u] = sb[k];
                        q[u] = sq[k];
                    }
                }
#pragma unroll
                for (int u = 0; u < 4; u++) {
                    const uint32_t k = k0 + 32u * u + lane;
                    if (k < ln) {
                        dd[k] = b[u];
                        dd[ln + 3u + k] = q[u];
                    }
                }
            }
        }
    }
}

// file_off[t][b] = first byte of bucket b's run in stream t (b = n_buckets: the end of the stream)
__global__ void k_emit_file_offsets(const unsigned long long* __restrict__ pre, const unsigned long long* __restrict__ bucket_off,
                                    const unsigned long long* __restrict__ stream_base, uint64_t n, uint32_t n_buckets,
                                    uint32_t n_streams, unsigned long long* __restrict__ file_off) {
    const uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n_streams * (n_buckets + 1)) return;
    const uint32_t t = k / (n_buckets + 1), b = k % (n_buckets + 1);
    file_off[k] = stream_base[t] + pre[(size_t)t * (n + 1) + bucket_off[b]];
}

int em_fail(int code, const std::string& msg) {
    fq::set_last_error(msg);
    return code;
}
#define EM_CU(call)                                                                                                  \
    do {                                                                                                             \
        cudaError_t e__ = (call);                                                                                    \
        if (e__ != cudaSuccess) return em_fail(FQTK_B200_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(e__)); \
    } while (0)

}  // namespace

extern "C" {

int fqtk_b200_emit_streams(const fqtk_b200_read_segment* segments, uint32_t n_segments, const char* output_kinds,
                           uint32_t* n_streams, char* stream_kinds, uint32_t* stream_numbers) {
    if (!segments || !output_kinds || !n_streams) return em_fail(FQTK_B200_ERR_ARG, "NULL argument");
    uint32_t ns = 0;
    for (const char* kind = "TBMC"; *kind; kind++) {  // the order SampleWriters::write walks its writers (demux.rs:397-402)
        if (!std::strchr(output_kinds, *kind)) continue;
        uint32_t idx = 0;
        for (uint32_t k = 0; k < n_segments; k++)
            if (segments[k].kind == (uint8_t)*kind) {
                if (ns >= EM_MAX_STREAMS) return em_fail(FQTK_B200_ERR_UNSUPPORTED, "more than 16 output streams");
                if (stream_kinds) stream_kinds[ns] = *kind;
                if (stream_numbers) stream_numbers[ns] = idx + 1;
                ns++;
                idx++;
            }
    }
    *n_streams = ns;
    return FQTK_B200_OK;
}

int fqtk_b200_demux_emit_device(int device, const fqtk_b200_emit_source* sources, uint32_t n_sources,
                                const fqtk_b200_read_segment* segments, uint32_t n_segments, const char* output_kinds,
                                const uint32_t* d_order, const uint64_t* d_offsets, uint32_t n_buckets, uint64_t n_reads,
                                uint8_t* d_text, uint64_t text_capacity, uint64_t* file_offsets, uint64_t* text_bytes,
                                void* stream) {
    if (!sources || !segments || !output_kinds || !file_offsets || !text_bytes || n_sources == 0 || n_sources > EM_MAX_SRC ||
        n_segments == 0 || n_segments > EM_MAX_SEG)
        return em_fail(FQTK_B200_ERR_ARG, "need 1..8 sources and 1..32 segments");
    EmitPlan p{};
    p.n_sources = n_sources;
    p.n_segs = n_segments;
    for (uint32_t s = 0; s < n_sources; s++) {
        if (!sources[s].d_chunk || !sources[s].d_seq_offsets || !sources[s].d_seq_lengths || (s == 0 && !sources[s].d_head_offsets))
            return em_fail(FQTK_B200_ERR_ARG, "NULL chunk / table pointer");
        p.chunk[s] = sources[s].d_chunk;
        p.head_off[s] = reinterpret_cast<const unsigned long long*>(sources[s].d_head_offsets);
        p.seq_off[s] = reinterpret_cast<const unsigned long long*>(sources[s].d_seq_offsets);
        p.seq_len[s] = sources[s].d_seq_lengths;
    }
    for (uint32_t k = 0; k < n_segments; k++) {
        if (segments[k].source >= n_sources) return em_fail(FQTK_B200_ERR_ARG, "segment names a source that was not given");
        if (!std::strchr("TBMSC", segments[k].kind) || segments[k].kind == 0) return em_fail(FQTK_B200_ERR_ARG, "segment kind must be one of T B M S C");
        p.seg[k] = EmitSeg{segments[k].source, segments[k].kind, segments[k].offset, segments[k].length};
    }
    uint32_t ns = 0;
    {
        uint32_t nums[EM_MAX_STREAMS];
        char kinds[EM_MAX_STREAMS];
        const int rc = fqtk_b200_emit_streams(segments, n_segments, output_kinds, &ns, kinds, nums);
        if (rc != FQTK_B200_OK) return rc;
        uint32_t t = 0;
        for (const char* kind = "TBMC"; *kind; kind++) {
            if (!std::strchr(output_kinds, *kind)) continue;
            for (uint32_t k = 0; k < n_segments; k++)
                if (segments[k].kind == (uint8_t)*kind) {
                    p.stream_seg[t] = k;
                    p.stream_readnum[t] = nums[t];
                    t++;
                }
        }
    }
    p.n_streams = ns;
    *text_bytes = 0;
    for (uint32_t k = 0; k < ns * (n_buckets + 1); k++) file_offsets[k] = 0;
    if (ns == 0 || n_reads == 0) return FQTK_B200_OK;
    if (!d_order || !d_offsets || !d_text) return em_fail(FQTK_B200_ERR_ARG, "NULL order / offsets / text buffer");
    if (n_reads >= (1ull << 32)) return em_fail(FQTK_B200_ERR_ARG, "n_reads must be < 2^32");
    EM_CU(cudaSetDevice(device));
    int sm = 0;
    EM_CU(cudaDeviceGetAttribute(&sm, cudaDevAttrMultiProcessorCount, device));
    cudaStream_t st = (cudaStream_t)stream;
    const uint64_t n = n_reads;
    const uint32_t n_tiles = (uint32_t)((n + EM_TILE - 1) / EM_TILE);
    fq::TempBuf lens, tiles, pre, base, err, foff;
    EM_CU(lens.alloc((size_t)ns * n * 4, st));
    EM_CU(tiles.alloc((size_t)ns * (n_tiles + 1) * 8, st));
    EM_CU(pre.alloc((size_t)ns * (n + 1) * 8, st));
    EM_CU(base.alloc((size_t)EM_MAX_STREAMS * 8, st));
    EM_CU(err.alloc(8, st));
    EM_CU(foff.alloc((size_t)ns * (n_buckets + 1) * 8, st));
    EM_CU(cudaMemsetAsync(err.p, 0xFF, 8, st));
    const uint32_t gx = (uint32_t)std::min<uint64_t>((n + EM_THREADS - 1) / EM_THREADS, (uint64_t)sm * 8);
    k_emit_lengths<<<dim3(gx, ns), EM_THREADS, 0, st>>>(p, d_order, n, static_cast<uint32_t*>(lens.p),
                                                       static_cast<unsigned long long*>(err.p));
    fq::count_launch();
    k_emit_tile_sums<<<dim3(n_tiles, ns), EM_THREADS, 0, st>>>(static_cast<uint32_t*>(lens.p), n, n_tiles,
                                                              static_cast<unsigned long long*>(tiles.p));
    fq::count_launch();
    k_emit_scan_tiles<<<ns, 1024, 0, st>>>(static_cast<unsigned long long*>(tiles.p), n_tiles);
    fq::count_launch();
    k_emit_apply<<<dim3(n_tiles, ns), EM_THREADS, 0, st>>>(static_cast<uint32_t*>(lens.p), n, n_tiles,
                                                          static_cast<unsigned long long*>(tiles.p),
                                                          static_cast<unsigned long long*>(pre.p));
    fq::count_launch();
    EM_CU(cudaGetLastError());
    unsigned long long h_err = 0;
    std::vector<unsigned long long> totals(ns), sbase(EM_MAX_STREAMS, 0);
    EM_CU(cudaMemcpyAsync(&h_err, err.p, 8, cudaMemcpyDeviceToHost, st));
    for (uint32_t t = 0; t < ns; t++)
        EM_CU(cudaMemcpyAsync(&totals[t], static_cast<unsigned long long*>(tiles.p) + (size_t)t * (n_tiles + 1) + n_tiles, 8,
                              cudaMemcpyDeviceToHost, st));
    EM_CU(cudaStreamSynchronize(st));
    if (h_err != ~0ull) {
        static const char* const what[4] = {"", "Can't handle read name with more than 8 segments", "Comment in did not have 4 segments",
                                            "empty comment after the read name"};
        return em_fail(FQTK_B200_ERR_ARG, std::string(what[h_err & 3u]) + " (read " + std::to_string(h_err >> 2) + ")");
    }
    unsigned long long total = 0;
    for (uint32_t t = 0; t < ns; t++) {
        sbase[t] = total;
        total += totals[t];
    }
    *text_bytes = total;
    if (total > text_capacity)
        return em_fail(FQTK_B200_ERR_ARG, "text buffer too small: " + std::to_string(total) + " bytes needed");
    EM_CU(cudaMemcpyAsync(base.p, sbase.data(), (size_t)EM_MAX_STREAMS * 8, cudaMemcpyHostToDevice, st));
    k_emit_write<<<dim3(gx, ns), EM_THREADS, 0, st>>>(p, d_order, n, static_cast<unsigned long long*>(pre.p),
                                                     static_cast<unsigned long long*>(base.p), d_text);
    fq::count_launch();
    const uint32_t nf = ns * (n_buckets + 1);
    k_emit_file_offsets<<<(nf + 255) / 256, 256, 0, st>>>(static_cast<unsigned long long*>(pre.p),
                                                         reinterpret_cast<const unsigned long long*>(d_offsets),
                                                         static_cast<unsigned long long*>(base.p), n, n_buckets, ns,
                                                         static_cast<unsigned long long*>(foff.p));
    fq::count_launch();
    EM_CU(cudaGetLastError());
    EM_CU(cudaMemcpyAsync(file_offsets, foff.p, (size_t)nf * 8, cudaMemcpyDeviceToHost, st));
    EM_CU(cudaStreamSynchronize(st));
    return FQTK_B200_OK;
}

}  // extern "C"
