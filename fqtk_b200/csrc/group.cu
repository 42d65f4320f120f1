// group.cu — the parts of the C ABI (include/fqtk_b200.h) that sit ABOVE a single matcher handle:
//   * fqtk_b200_group_*     one process, several GPUs of one box: one matcher per device, a batch split
//                           into contiguous shards with one host thread per device, and the per-sample count table
//                           summed on the first device over NVLink peer access (the reference keeps ONE table:
//                           src/bin/commands/demux.rs:921-926, 970-975)
//   * fqtk_b200_pack_host   encode() (src/lib/mod.rs:49-61) for a batch on the host: two symbols per table lookup, or the
//                           AVX2 stream form of host_pack.cpp when the rows lie back to back and L is a multiple of 8
//   * fqtk_b200_copy_ceiling   the pinned-memory copy rate of the platform (no kernels): the ceiling any host-buffer call
//                           is measured against
// Everything here goes through the public single-matcher entry points; it holds no matching logic of its own.
#include <atomic>
#include <chrono>
#include <cstdlib>
#include <cstring>
#include <new>
#include <string>
#include <thread>
#include <vector>

#include "../../include/fqtk_b200.h"
#include "common.cuh"

namespace fq {
void count_launch();
void set_last_error(const std::string& msg);
bool have_avx2();                                                                                 // host_pack.cpp
void pack_stream(const uint8_t* in, uint64_t n_bytes, uint8_t* out, const uint8_t* lut);  // host_pack.cpp
}  // namespace fq

namespace {

int g_fail(int code, const std::string& msg) {
    fq::set_last_error(msg);
    return code;
}

// counts[0][b] += counts[k][b] for every peer table k >= 1: peers are read straight through their mapped pointers
// (NVLink P2P loads) or, where peer access is not available, from copies staged on this device
__global__ void k_sum_count_tables(unsigned long long* __restrict__ total, const unsigned long long* const* __restrict__ tables,
                                   uint32_t n_tables, uint32_t n_bins) {
    for (uint32_t b = blockIdx.x * blockDim.x + threadIdx.x; b < n_bins; b += gridDim.x * blockDim.x) {
        unsigned long long v = 0;
        for (uint32_t k = 0; k < n_tables; k++) v += tables[k][b];
        total[b] = v;
    }
}

}  // namespace

struct fqtk_b200_group {
    std::vector<fqtk_b200_matcher*> matchers;
    std::vector<int> devices;
    uint32_t S = 0;
    // on devices[0]: the summed table, the pointer list the sum kernel walks, staging for peers without P2P access
    unsigned long long* d_total = nullptr;
    const unsigned long long** d_tables = nullptr;
    unsigned long long* d_stage = nullptr;
    std::vector<char> peer_ok;  // peer access from devices[0] to devices[k]
    cudaStream_t stream = nullptr;
};

namespace {
// run `call(k, first, count)` for every device's shard on its own host thread; first error wins
template <typename F>
int for_each_shard(fqtk_b200_group* g, uint64_t n, F call) {
    const size_t G = g->matchers.size();
    std::vector<int> rcs(G, FQTK_B200_OK);
    std::vector<std::string> errs(G);
    std::vector<std::thread> threads;
    for (size_t k = 0; k < G; k++)
        threads.emplace_back([&, k] {
            uint64_t first = 0, count = 0;
            fqtk_b200_group_shard(g, n, (uint32_t)k, &first, &count);
            if (count == 0) return;
            rcs[k] = call(k, first, count);
            if (rcs[k] != FQTK_B200_OK) errs[k] = fqtk_b200_last_error();
        });
    for (auto& t : threads) t.join();
    for (size_t k = 0; k < G; k++)
        if (rcs[k] != FQTK_B200_OK) return g_fail(rcs[k], errs[k]);
    return FQTK_B200_OK;
}
}  // namespace

extern "C" {

int fqtk_b200_group_create(const uint8_t* panel_ascii, uint32_t S, uint32_t L, uint8_t max_mm, uint8_t min_delta,
                           int use_cache, const int* devices, uint32_t n_devices, const fqtk_b200_options* opts,
                           fqtk_b200_group** out) {
    if (!out) return g_fail(FQTK_B200_ERR_ARG, "out is NULL");
    *out = nullptr;
    int ndev = fqtk_b200_device_count();
    if (ndev == 0) return g_fail(FQTK_B200_ERR_CUDA, "no CUDA device: fqtk_b200 has no CPU fallback");
    std::vector<int> devs;
    if (devices == nullptr || n_devices == 0) {  // every visible device
        for (int d = 0; d < ndev; d++) devs.push_back(d);
    } else {
        for (uint32_t k = 0; k < n_devices; k++) {
            if (devices[k] < 0 || devices[k] >= ndev) return g_fail(FQTK_B200_ERR_ARG, "bad device ordinal");
            devs.push_back(devices[k]);  // (a device may be listed more than once: two handles on one GPU, for testing)
        }
    }
    fqtk_b200_group* g = new (std::nothrow) fqtk_b200_group();
    if (!g) return g_fail(FQTK_B200_ERR_ARG, "out of host memory");
    g->devices = devs;
    g->S = S;
    g->matchers.assign(devs.size(), nullptr);
    // one host thread per device: the panel's tables are built concurrently (nothing about a handle is process-global)
    std::vector<int> rcs(devs.size(), FQTK_B200_OK);
    std::vector<std::string> errs(devs.size());
    std::vector<std::thread> threads;
    for (size_t k = 0; k < devs.size(); k++)
        threads.emplace_back([&, k] {
            rcs[k] = fqtk_b200_matcher_create_ex(panel_ascii, S, L, max_mm, min_delta, use_cache, devs[k], opts,
                                                 &g->matchers[k]);
            if (rcs[k] != FQTK_B200_OK) errs[k] = fqtk_b200_last_error();
        });
    for (auto& t : threads) t.join();
    for (size_t k = 0; k < devs.size(); k++)
        if (rcs[k] != FQTK_B200_OK) {
            const int rc = rcs[k];
            const std::string msg = errs[k];
            fqtk_b200_group_destroy(g);
            return g_fail(rc, msg);
        }
    // the count reduction lives on the first device
    cudaError_t e = cudaSetDevice(devs[0]);
    if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&g->stream, cudaStreamNonBlocking);
    if (e == cudaSuccess) e = cudaMalloc(&g->d_total, (size_t)(S + 1) * 8);
    if (e == cudaSuccess) e = cudaMalloc(&g->d_tables, devs.size() * sizeof(void*));
    if (e == cudaSuccess) e = cudaMalloc(&g->d_stage, devs.size() * (size_t)(S + 1) * 8);
    g->peer_ok.assign(devs.size(), 0);
    std::vector<const unsigned long long*> tables(devs.size());
    for (size_t k = 0; k < devs.size() && e == cudaSuccess; k++) {
        uint64_t* dc = nullptr;
        fqtk_b200_matcher_counts_device(g->matchers[k], &dc);
        int can = (k == 0);
        if (k > 0 && devs[k] == devs[0]) can = 1;  // the same device listed again: its memory is simply local
        if (k > 0 && devs[k] != devs[0]) {
            cudaDeviceCanAccessPeer(&can, devs[0], devs[k]);
            if (can) {
                const cudaError_t pe = cudaDeviceEnablePeerAccess(devs[k], 0);
                if (pe == cudaErrorPeerAccessAlreadyEnabled) cudaGetLastError();
                else if (pe != cudaSuccess) { can = 0; cudaGetLastError(); }
            }
        }
        g->peer_ok[k] = (char)can;
        tables[k] = can ? reinterpret_cast<const unsigned long long*>(dc) : g->d_stage + k * (size_t)(S + 1);
    }
    if (e == cudaSuccess) e = cudaMemcpy(g->d_tables, tables.data(), devs.size() * sizeof(void*), cudaMemcpyHostToDevice);
    if (e != cudaSuccess) {
        fqtk_b200_group_destroy(g);
        return g_fail(FQTK_B200_ERR_CUDA, std::string("group_create: ") + cudaGetErrorString(e));
    }
    *out = g;
    return FQTK_B200_OK;
}

void fqtk_b200_group_destroy(fqtk_b200_group* g) {
    if (!g) return;
    for (fqtk_b200_matcher* m : g->matchers) fqtk_b200_matcher_destroy(m);
    if (!g->devices.empty()) cudaSetDevice(g->devices[0]);
    if (g->d_total) cudaFree(g->d_total);
    if (g->d_tables) cudaFree(g->d_tables);
    if (g->d_stage) cudaFree(g->d_stage);
    if (g->stream) cudaStreamDestroy(g->stream);
    delete g;
}

uint32_t fqtk_b200_group_size(const fqtk_b200_group* g) { return g ? (uint32_t)g->matchers.size() : 0u; }

fqtk_b200_matcher* fqtk_b200_group_matcher(fqtk_b200_group* g, uint32_t k) {
    return (g && k < g->matchers.size()) ? g->matchers[k] : nullptr;
}

int fqtk_b200_group_device(const fqtk_b200_group* g, uint32_t k) {
    return (g && k < g->devices.size()) ? g->devices[k] : -1;
}

void fqtk_b200_group_shard(const fqtk_b200_group* g, uint64_t n_reads, uint32_t k, uint64_t* first, uint64_t* count) {
    const uint64_t G = g ? g->matchers.size() : 1;
    // contiguous, tiles [0, n) exactly, sizes differ by at most one read; starts rounded to 4 reads so that a packed
    // shard begins on a 16-byte boundary whatever W is
    auto bound = [&](uint64_t i) { return i >= G ? n_reads : ((n_reads / G * i + (n_reads % G) * i / G) & ~3ull); };
    if (first) *first = bound(k);
    if (count) *count = bound(k + 1) - bound(k);
}

int fqtk_b200_group_assign_batch(fqtk_b200_group* g, const uint8_t* rows, uint64_t n, uint64_t stride,
                                 const uint32_t* lengths, uint32_t* results) {
    if (!g || (n && (!rows || !results))) return g_fail(FQTK_B200_ERR_ARG, "NULL argument");
    return for_each_shard(g, n, [&](size_t k, uint64_t first, uint64_t count) {
        return fqtk_b200_matcher_assign_batch(g->matchers[k], rows + first * stride, count, stride,
                                              lengths ? lengths + first : nullptr, results + first);
    });
}

int fqtk_b200_group_assign_batch_packed(fqtk_b200_group* g, const uint32_t* packed, uint64_t n, uint32_t* results,
                                        uint16_t* sample_index) {
    if (!g || (n && !packed) || (n && !results && !sample_index)) return g_fail(FQTK_B200_ERR_ARG, "NULL argument");
    fqtk_b200_matcher_info info{};
    fqtk_b200_matcher_get_info(g->matchers[0], &info);
    const uint64_t W = info.words_per_read;
    return for_each_shard(g, n, [&](size_t k, uint64_t first, uint64_t count) {
        return fqtk_b200_matcher_assign_batch_packed(g->matchers[k], packed + first * W, count,
                                                     results ? results + first : nullptr,
                                                     sample_index ? sample_index + first : nullptr);
    });
}

int fqtk_b200_group_assign_packed_device(fqtk_b200_group* g, const uint32_t* const* d_packed, const uint64_t* n_reads,
                                         uint32_t* const* d_results, void* const* streams) {
    if (!g || !d_packed || !n_reads || !d_results) return g_fail(FQTK_B200_ERR_ARG, "NULL argument");
    for (size_t k = 0; k < g->matchers.size(); k++) {  // asynchronous launches: no host threads needed
        const int rc = fqtk_b200_matcher_assign_packed_device(g->matchers[k], d_packed[k], n_reads[k], d_results[k],
                                                              streams ? streams[k] : nullptr);
        if (rc != FQTK_B200_OK) return rc;
    }
    return FQTK_B200_OK;
}

int fqtk_b200_group_counts(fqtk_b200_group* g, uint64_t* out) {
    if (!g || !out) return g_fail(FQTK_B200_ERR_ARG, "NULL argument");
    const size_t G = g->matchers.size();
    const size_t bytes = (size_t)(g->S + 1) * 8;
    cudaError_t e = cudaSuccess;
    for (size_t k = 0; k < G && e == cudaSuccess; k++) {  // everything the devices were asked to do is done
        e = cudaSetDevice(g->devices[k]);
        if (e == cudaSuccess) e = cudaDeviceSynchronize();
    }
    if (e == cudaSuccess) e = cudaSetDevice(g->devices[0]);
    for (size_t k = 1; k < G && e == cudaSuccess; k++)
        if (!g->peer_ok[k]) {  // no peer access: stage that device's table here first
            uint64_t* dc = nullptr;
            fqtk_b200_matcher_counts_device(g->matchers[k], &dc);
            e = cudaMemcpyPeerAsync(g->d_stage + k * (size_t)(g->S + 1), g->devices[0], dc, g->devices[k], bytes, g->stream);
        }
    if (e == cudaSuccess) {
        k_sum_count_tables<<<(g->S + 1 + 255) / 256, 256, 0, g->stream>>>(g->d_total, g->d_tables, (uint32_t)G, g->S + 1);
        fq::count_launch();
        e = cudaGetLastError();
    }
    if (e == cudaSuccess) e = cudaMemcpyAsync(out, g->d_total, bytes, cudaMemcpyDeviceToHost, g->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(g->stream);
    if (e != cudaSuccess) return g_fail(FQTK_B200_ERR_CUDA, std::string("group_counts: ") + cudaGetErrorString(e));
    return FQTK_B200_OK;
}

int fqtk_b200_group_reset_counts(fqtk_b200_group* g) {
    if (!g) return g_fail(FQTK_B200_ERR_ARG, "NULL group");
    for (fqtk_b200_matcher* m : g->matchers) {
        const int rc = fqtk_b200_matcher_reset_counts(m);
        if (rc != FQTK_B200_OK) return rc;
    }
    return FQTK_B200_OK;
}

// ---- encode() for a batch on the host -----------------------------------------------------------------------------
int fqtk_b200_pack_host(const uint8_t* rows, uint64_t n, uint32_t L, uint64_t stride, uint32_t* out_packed, int threads) {
    if ((n && (!rows || !out_packed)) || L == 0 || L > FQTK_B200_MAX_BARCODE_LEN || stride < L)
        return g_fail(FQTK_B200_ERR_ARG, "bad argument");
    // two symbols per lookup: byte pair -> two 4-bit masks (64 KiB table, built once; thread-safe static initialisation)
    static const std::vector<uint8_t> lut2 = [] {
        std::vector<uint8_t> t(65536);
        for (uint32_t v = 0; v < 65536u; v++)
            t[v] = (uint8_t)(fq::encode_byte(v & 0xFFu) | (fq::encode_byte(v >> 8) << 4));
        return t;
    }();
    static const std::vector<uint8_t> lut1 = [] {
        std::vector<uint8_t> t(256);
        for (uint32_t v = 0; v < 256u; v++) t[v] = (uint8_t)fq::encode_byte(v);
        return t;
    }();
    const uint32_t W = fq::words_for_len(L);
    const uint8_t* lut = lut2.data();
    // rows back to back and L a multiple of 8 (cfg 2, 3, 4): the batch is one stream of symbols -> host_pack.cpp (AVX2)
    const bool stream = stride == L && (L % 8u) == 0u && fq::have_avx2();
    auto work = [&](uint64_t lo, uint64_t hi) {
        if (stream) {
            fq::pack_stream(rows + lo * L, (hi - lo) * L, reinterpret_cast<uint8_t*>(out_packed + lo * W), lut1.data());
            return;
        }
        for (uint64_t i = lo; i < hi; i++) {
            const uint8_t* r = rows + i * stride;
            uint32_t* o = out_packed + i * W;
            uint32_t k = 0;
            for (uint32_t w = 0; w < W; w++) {
                uint32_t acc = 0;
                uint32_t sh = 0;
                for (; sh < 32u && k + 1u < L; sh += 8u, k += 2u) {
                    uint16_t pair;
                    std::memcpy(&pair, r + k, 2);  // little-endian: first symbol in the low byte
                    acc |= (uint32_t)lut[pair] << sh;
                }
                if (sh < 32u && k < L) {  // odd tail symbol
                    acc |= fq::encode_byte(r[k]) << sh;
                    k++;
                }
                o[w] = acc;
            }
        }
    };
    int T = threads > 0 ? threads : (int)std::thread::hardware_concurrency();
    T = (int)std::max<uint64_t>(1, std::min<uint64_t>((uint64_t)std::max(T, 1), n / 65536 + 1));
    if (T == 1) {
        work(0, n);
        return FQTK_B200_OK;
    }
    std::vector<std::thread> pool;
    for (int t = 0; t < T; t++)  // (ranges start at multiples of four rows: 16-byte aligned words for host_pack.cpp)
        pool.emplace_back(work, (n * t / T) & ~3ull, t + 1 == T ? n : ((n * (t + 1) / T) & ~3ull));
    for (auto& t : pool) t.join();
    return FQTK_B200_OK;
}

// ---- 4-line FASTQ record scanner (the part of seq_io's reader the GPU path needs: where the sequence lines are) ----
int fqtk_b200_fastq_scan(const uint8_t* chunk, uint64_t chunk_bytes, uint64_t max_records, uint64_t* head_offsets,
                         uint64_t* seq_offsets, uint32_t* seq_lengths, uint64_t* n_records, uint64_t* consumed) {
    if ((chunk_bytes && !chunk) || !seq_offsets || !seq_lengths || !n_records || !consumed)
        return g_fail(FQTK_B200_ERR_ARG, "NULL argument");
    uint64_t pos = 0, n = 0;
    *n_records = 0;
    *consumed = 0;
    while (n < max_records && pos < chunk_bytes) {
        uint64_t start[4], len[4];
        uint64_t p = pos;
        bool whole = true;
        for (int line = 0; line < 4; line++) {
            const void* nl = p < chunk_bytes ? std::memchr(chunk + p, '\n', (size_t)(chunk_bytes - p)) : nullptr;
            if (!nl) {
                whole = false;  // the record runs past the chunk: the caller carries it over
                break;
            }
            const uint64_t end = (uint64_t)(static_cast<const uint8_t*>(nl) - chunk);
            start[line] = p;
            len[line] = end - p;
            if (len[line] && chunk[end - 1] == '\r') len[line]--;
            p = end + 1;
        }
        if (!whole) break;
        if (len[0] == 0 || chunk[start[0]] != '@')
            return g_fail(FQTK_B200_ERR_ARG, "FASTQ record " + std::to_string(n) + ": header line does not start with '@'");
        if (len[2] == 0 || chunk[start[2]] != '+')
            return g_fail(FQTK_B200_ERR_ARG, "FASTQ record " + std::to_string(n) + ": separator line does not start with '+'");
        if (len[1] != len[3])
            return g_fail(FQTK_B200_ERR_ARG, "FASTQ record " + std::to_string(n) + ": sequence and quality lengths differ");
        if (len[1] > 0xFFFFFFFFull) return g_fail(FQTK_B200_ERR_ARG, "FASTQ record too long");
        if (head_offsets) head_offsets[n] = start[0];
        seq_offsets[n] = start[1];
        seq_lengths[n] = (uint32_t)len[1];
        n++;
        pos = p;
    }
    *n_records = n;
    *consumed = pos;
    return FQTK_B200_OK;
}

// ---- the platform's pinned-memory copy ceiling ----------------------------------------------------------------------
int fqtk_b200_copy_ceiling(int device, uint64_t in_bytes, uint64_t out_bytes, uint64_t chunk_in_bytes, int reps,
                           double* seconds_per_rep) {
    if (!seconds_per_rep || reps < 1 || in_bytes == 0) return g_fail(FQTK_B200_ERR_ARG, "bad argument");
    if (chunk_in_bytes == 0) chunk_in_bytes = 32ull << 20;
    cudaError_t e = cudaSetDevice(device);
    if (e != cudaSuccess) return g_fail(FQTK_B200_ERR_CUDA, cudaGetErrorString(e));
    const uint64_t n_chunks = (in_bytes + chunk_in_bytes - 1) / chunk_in_bytes;
    const uint64_t chunk_out = (out_bytes + n_chunks - 1) / n_chunks;
    constexpr int NP = 3;
    uint8_t *h_in = nullptr, *h_out = nullptr, *d_in[NP] = {}, *d_out[NP] = {};
    cudaStream_t st[NP] = {};
    e = cudaHostAlloc(&h_in, in_bytes, cudaHostAllocPortable);
    if (e == cudaSuccess) e = cudaHostAlloc(&h_out, std::max<uint64_t>(out_bytes, 1), cudaHostAllocPortable);
    for (int s = 0; s < NP && e == cudaSuccess; s++) {
        e = cudaMalloc(&d_in[s], chunk_in_bytes);
        if (e == cudaSuccess) e = cudaMalloc(&d_out[s], std::max<uint64_t>(chunk_out, 1));
        if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&st[s], cudaStreamNonBlocking);
    }
    if (e == cudaSuccess) {
        std::memset(h_in, 0x41, in_bytes);  // first touch on this thread's NUMA node
        std::memset(h_out, 0, std::max<uint64_t>(out_bytes, 1));
    }
    double best = 0;
    for (int rep = 0; rep <= reps && e == cudaSuccess; rep++) {  // rep 0 warms up
        const auto t0 = std::chrono::steady_clock::now();
        uint64_t in_done = 0, out_done = 0;
        for (uint64_t c = 0; c < n_chunks && e == cudaSuccess; c++) {
            const int s = (int)(c % NP);
            const uint64_t ib = std::min(chunk_in_bytes, in_bytes - in_done);
            const uint64_t ob = std::min(chunk_out, out_bytes - out_done);
            e = cudaMemcpyAsync(d_in[s], h_in + in_done, ib, cudaMemcpyHostToDevice, st[s]);
            if (e == cudaSuccess && ob) e = cudaMemcpyAsync(h_out + out_done, d_out[s], ob, cudaMemcpyDeviceToHost, st[s]);
            in_done += ib;
            out_done += ob;
        }
        for (int s = 0; s < NP && e == cudaSuccess; s++) e = cudaStreamSynchronize(st[s]);
        const double dt = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
        if (rep > 0) best += dt;
    }
    for (int s = 0; s < NP; s++) {
        if (d_in[s]) cudaFree(d_in[s]);
        if (d_out[s]) cudaFree(d_out[s]);
        if (st[s]) cudaStreamDestroy(st[s]);
    }
    if (h_in) cudaFreeHost(h_in);
    if (h_out) cudaFreeHost(h_out);
    if (e != cudaSuccess) return g_fail(FQTK_B200_ERR_CUDA, std::string("copy_ceiling: ") + cudaGetErrorString(e));
    *seconds_per_rep = best / reps;
    return FQTK_B200_OK;
}

}  // extern "C"
