// C ABI of libnanosim_b200.so (include/nanosim_b200.h): context, HBM residency of reference + model tables,
// batch orchestration (plan -> scan -> script -> emit) on one CUDA stream, device->host fetch.
#include <cuda_runtime.h>
#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_scan.cuh>

#include <algorithm>
#include <chrono>
#if defined(__x86_64__)
#include <immintrin.h>
#endif
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <dlfcn.h>
#include <fcntl.h>
#include <sched.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>
#include <string>
#include <mutex>
#include <thread>
#include <vector>

#if __has_include(<nccl.h>)
#include <nccl.h>
#define NS_HAVE_NCCL 1
#else
#define NS_HAVE_NCCL 0
#endif

#include "nanosim_b200.h"
#include "device_common.cuh"
#include "plan_kernel.cuh"
#include "emit_kernel.cuh"
#include "uread_kernel.cuh"
#include "hp_kernel.cuh"

namespace {

struct DevBuf {
    void* p = nullptr;
    size_t cap = 0;
    cudaError_t ensure(size_t bytes) {
        if (bytes <= cap) return cudaSuccess;
        if (p) cudaFree(p);
        p = nullptr;
        cap = 0;
        size_t want = bytes + bytes / 8 + 256;
        cudaError_t e = cudaMalloc(&p, want);
        if (e == cudaSuccess) cap = want;
        return e;
    }
    // grow but keep the first `keep` bytes (rare path)
    cudaError_t ensure_keep(size_t bytes, size_t keep, cudaStream_t st) {
        if (bytes <= cap) return cudaSuccess;
        void* np = nullptr;
        size_t want = bytes + bytes / 8 + 256;
        cudaError_t e = cudaMalloc(&np, want);
        if (e != cudaSuccess) return e;
        if (p && keep) {
            e = cudaMemcpyAsync(np, p, keep, cudaMemcpyDeviceToDevice, st);
            if (e == cudaSuccess) e = cudaStreamSynchronize(st);
        }
        if (p) cudaFree(p);
        p = np;
        cap = want;
        return e;
    }
    void release() {
        if (p) cudaFree(p);
        p = nullptr;
        cap = 0;
    }
    template <class T> T* as() const { return reinterpret_cast<T*>(p); }
};

}  // namespace

struct NsContext {
    int device = 0;
    uint64_t seed = 0;
    cudaStream_t stream = nullptr;
    cudaEvent_t ev[6] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
    std::string err;
    int sm_count = 148;

    bool borrowed = false;          // ns_clone: reference + model buffers belong to the parent
    bool have_ref = false, have_model = false, have_cfg = false;
    DevBuf ref_bases, ref_off, ref_packed, ref_pk_off, ref_exc;
    DevBuf trx_sorted;              // transcript lengths ascending + their records (ns_configure, transcriptome mode)
    bool trx_sorted_owned = false;  // false on a clone that still uses its parent's table
    DevRef dref{};
    std::vector<uint64_t> h_chrom_off;

    DevBuf kde[5], alias, qlut, ref_species, ref_circular, ref_sp_off, kde2d_x, kde2d_y, expr_alias, expr_chrom, chrom_polya;
    bool have_expr = false;
    std::vector<uint32_t> h_sp_off;
    std::vector<double> abun, abun_inflated, species_bases;     // metagenome: dict_abun, dict_abun_inflated, running totals
    DevBuf sp_bases_dev;
    DevModel dmodel{};
    NsModel hmodel{};
    NsRunConfig hcfg{};
    DevCfg dcfg{};

    // batch state
    DevBuf split_base, split_extra, split_ckpt, hp_keys;     // long pieces -> extra emit work items (emit_kernel.cuh:split_kernel)
    DevBuf reads, pieces, ops, seq, qual, nseg, npieces, piece_first, scan_in, scan_out, scan_tmp, counter, totals,
        stats, sort_keys, sort_vals, sort_tmp, hp_off;
    uint64_t* h_totals = nullptr;   // pinned + mapped
    uint64_t* h_totals_dev = nullptr;
    // ns_fetch sends bases over PCIe as 2 bits each: device pack buffer, pinned staging, event after its copy
    DevBuf pack_dev;
    uint8_t* pack_host = nullptr;
    size_t pack_host_cap = 0;
    cudaEvent_t ev_pack = nullptr;
    cudaEvent_t ev_block = nullptr;     // cudaEventBlockingSync: long waits sleep instead of spinning (wait_stream)
    // Batches of one job have near-identical sizes: once a batch of a kind has run with host-sized buffers, the next ones
    // are submitted in one go (no host round trip between the first and the last kernel) against those capacities; a
    // kernel checks them on the device and a batch that does not fit is simply run again the sized way.
    bool opt_ok[2] = {false, false};
    uint32_t opt_n[2] = {0, 0};
    NsBatchInfo last{};
    int last_kind = 0;
    uint64_t last_first_id = 0;
    bool have_batch = false;
};

// Wait for everything queued on the context's stream.  cudaStreamSynchronize spins: a host thread per overlapped context then
// burns a core for the whole 20-100 ms of a big batch.  The boxes of this pool grant their GPU processes a CPU quota (16 cores
// for 1 GPU, 24 for 2, 96 logical for 8), and per-GPU end-to-end throughput fell with the number of GPU processes (39 / 30 /
// 16 Gbases/s) although no PCIe link was saturated and the expansion of the 2-bit bases takes 19 ms of a 100 ms fetch: four
// spinning threads per process are the largest CPU consumer left.  With several GPU processes on the host (torchrun's
// LOCAL_WORLD_SIZE > 1, or NANOSIM_B200_BLOCKING_SYNC=1) long waits therefore sleep on a cudaEventBlockingSync event; a single
// process, and every short wait, spins as before.  (Written after the round's GPU minutes were spent: not measured.)
static bool blocking_sync_wanted() {
    static const bool want = [] {
        if (const char* e = getenv("NANOSIM_B200_BLOCKING_SYNC")) return atoi(e) != 0;
        const char* lw = getenv("LOCAL_WORLD_SIZE");
        return lw && atoi(lw) > 1;
    }();
    return want;
}
static cudaError_t wait_stream(NsContext* ctx, bool long_wait) {
    if (!long_wait || !blocking_sync_wanted()) return cudaStreamSynchronize(ctx->stream);
    cudaError_t e = cudaSuccess;
    if (!ctx->ev_block) e = cudaEventCreateWithFlags(&ctx->ev_block, cudaEventBlockingSync | cudaEventDisableTiming);
    if (e == cudaSuccess) e = cudaEventRecord(ctx->ev_block, ctx->stream);
    if (e == cudaSuccess) e = cudaEventSynchronize(ctx->ev_block);
    return e;
}


namespace {

cudaEvent_t g_base[64] = {};      // per device: origin of the device timeline reported in NsBatchInfo

int env_int(const char* name, int dflt) {
    const char* e = getenv(name);
    return (e && *e) ? atoi(e) : dflt;
}

int fail(NsContext* c, int code, const char* fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    if (c) c->err = buf;
    return code;
}

#define CK(call)                                                                                      \
    do {                                                                                              \
        cudaError_t e_ = (call);                                                                      \
        if (e_ != cudaSuccess)                                                                        \
            return fail(ctx, NS_ECUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); \
    } while (0)

__global__ void scatter_piece_off(NsPieceMeta* pieces, uint32_t n, const uint64_t* off) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) pieces[i].op_off = off[i];
}
__global__ void set_sentinel_off(NsPieceMeta* pieces, uint32_t n, const uint64_t* total) {
    if (threadIdx.x == 0 && blockIdx.x == 0) pieces[n].op_off = *total;
}
// pieces of reads whose script overflowed its slot (NsReadMeta.flags bit 0) get exact offsets behind the primary area
__global__ void gather_flagged_ops(const NsPieceMeta* pieces, const NsReadMeta* reads, uint32_t n, uint64_t* out) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = (reads[pieces[i].read_slot].flags & 1u) ? pieces[i].n_ops : 0u;
}
__global__ void scatter_flagged_off(NsPieceMeta* pieces, const NsReadMeta* reads, uint32_t n, const uint64_t* off, uint64_t base) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n && (reads[pieces[i].read_slot].flags & 1u)) pieces[i].op_off = base + off[i];
}
__global__ void copy_ev_fields(NsPieceMeta* pieces, uint32_t n) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) {
        pieces[i].ev_off = pieces[i].op_off;
        pieces[i].ev_n_ops = pieces[i].n_ops;
    }
}
__global__ void add_base_u64(uint64_t* v, uint32_t n, uint64_t base) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) v[i] += base;
}
// ns_set_reference: the 2-bit copy of the reference (DevRef::packed).  One thread per packed word; exceptions (bytes
// that are not a/c/g/t in either case) are counted per 256-word block, other[0] counts bytes that are not even an IUPAC
// nucleotide code (case_convert passes those through unchanged, so reads may then hold characters other than ACGT).
__global__ void pack_reference_kernel(const uint8_t* __restrict__ bases, const uint64_t* __restrict__ chrom_off,
                                      const uint64_t* __restrict__ pk_off, uint32_t n_chrom, uint32_t* __restrict__ packed,
                                      uint64_t n_words, uint32_t* __restrict__ exc_cnt, unsigned long long* other) {
    const uint64_t w = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (w >= n_words) return;
    // chromosome of this word: last c with pk_off[c] <= w
    uint32_t lo = 0, hi = n_chrom;
    while (hi - lo > 1) {
        const uint32_t mid = (lo + hi) >> 1;
        if (pk_off[mid] <= w) lo = mid; else hi = mid;
    }
    const uint64_t cstart = chrom_off[lo], clen = chrom_off[lo + 1] - cstart;
    const uint64_t first = (w - pk_off[lo]) * 16;
    uint32_t word = 0, n_exc = 0, n_other = 0;
    for (uint32_t j = 0; j < 16; ++j) {
        if (first + j >= clen) break;
        uint32_t c = bases[cstart + first + j];
        if (c - 'a' < 26u) c -= 32;
        if (acgt_fast(c)) {
            word |= base_idx(c) << (2 * j);
        } else {
            ++n_exc;
            if (resolve_iupac(c, 0u, 0u) == c) ++n_other;      // not in case_convert's table: passes through unchanged
        }
    }
    packed[w] = word;
    if (n_exc) atomicAdd(&exc_cnt[w >> REF_EXC_BLOCK_SHIFT], n_exc);
    if (n_other) atomicAdd(other, (unsigned long long)n_other);
}

// Bases leave the device as 2 bits each (the emit kernel only writes A C G T/U): 16 ASCII bytes -> one 32-bit word,
// base j of a byte quadruple in bits [2j, 2j+1], code = (c >> 1) & 3 (A 0, C 1, T/U 2, G 3).  ns_fetch expands them again
// on the host, so callers see the same ASCII buffers while the PCIe transfer of a FASTQ batch shrinks from 2 to 1.25 B/base.
__global__ void pack_bases_kernel(const uint4* __restrict__ seq, uint32_t* __restrict__ out, uint64_t n16) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n16) return;
    const uint4 v = seq[i];
    auto pk = [](uint32_t w) {
        const uint32_t t = (w >> 1) & 0x03030303u;
        return (t | (t >> 6) | (t >> 12) | (t >> 18)) & 0xffu;
    };
    out[i] = pk(v.x) | (pk(v.y) << 8) | (pk(v.z) << 16) | (pk(v.w) << 24);
}
// The batch totals the host needs mid-pipeline are written straight into mapped pinned host memory: a cudaMemcpy
// would queue behind another context's multi-GB device->host transfer on the copy engine and serialise the pipelines.
__global__ void publish_totals(const uint64_t* totals, volatile uint64_t* host) {
    if (threadIdx.x < 16) host[threadIdx.x] = totals[threadIdx.x];
    __threadfence_system();
}
// totals[] slots: 0 pieces, 1 overflow ops, 2 sequence bytes, 3 bases, 4 slot ops, 5 flagged reads, 6 hp ops, 7 pool cursor,
// 8 pool base, 9 pool size, 10 primary ops, 11 abort flags (bit 0: script area too small, bit 1: with the overflow area,
// bit 2: sequence buffers too small)
#define NS_T_POOL 8
#define NS_T_PRIMARY 10
#define NS_T_ABORT 11
// sync-free batches: what the host would compute from the scans, computed and checked against the capacities on the device
__global__ void capacity_stage_a(uint64_t* totals, uint32_t fast_unaligned, uint64_t ops_cap) {
    if (threadIdx.x || blockIdx.x) return;
    const uint64_t slot_ops = totals[4];
    const uint64_t pool_ops = fast_unaligned ? slot_ops / 16 + (1u << 20) : 0;
    totals[NS_T_POOL] = slot_ops;
    totals[NS_T_POOL + 1] = pool_ops;
    totals[NS_T_PRIMARY] = slot_ops + pool_ops;
    if (slot_ops + pool_ops + 4 > ops_cap) totals[NS_T_ABORT] |= 1u;
}
__global__ void capacity_stage_b(uint64_t* totals, uint64_t ops_cap, uint64_t seq_cap) {
    if (threadIdx.x || blockIdx.x) return;
    if (totals[NS_T_PRIMARY] + totals[1] + 4 > ops_cap) totals[NS_T_ABORT] |= 2u;
    if (totals[2] + 16 > seq_cap) totals[NS_T_ABORT] |= 4u;
}
__global__ void scatter_flagged_off_dev(NsPieceMeta* pieces, const NsReadMeta* reads, uint32_t n, const uint64_t* off, const uint64_t* base) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n && (reads[pieces[i].read_slot].flags & 1u)) pieces[i].op_off = *base + off[i];
}
__global__ void gather_read_bytes(const NsReadMeta* reads, uint32_t n, uint64_t* out) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = ((uint64_t)reads[i].seq_len + 15u) & ~(uint64_t)15u;
}
__global__ void scatter_read_off(NsReadMeta* reads, uint32_t n, const uint64_t* off) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) reads[i].seq_off = off[i];
}
__global__ void widen_u32(const uint32_t* in, uint32_t n, uint64_t* out) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = in[i];
}
__global__ void narrow_u64(const uint64_t* in, uint32_t n, uint32_t* out) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = (uint32_t)in[i];
}
// totals[k] = off[n-1] + in[n-1]
__global__ void last_total(const uint64_t* in, const uint64_t* off, uint32_t n, uint64_t* totals, int k) {
    if (threadIdx.x == 0 && blockIdx.x == 0) totals[k] = n ? off[n - 1] + in[n - 1] : 0;
}
__global__ void sum_bases(const NsReadMeta* reads, uint32_t n, unsigned long long* out) {
    unsigned long long s = 0;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) s += reads[i].seq_len;
    for (int d = 16; d > 0; d >>= 1) s += __shfl_down_sync(0xffffffffu, s, d);
    if ((threadIdx.x & 31) == 0 && s) atomicAdd(out, s);
}

// op-list histograms (ns_op_stats)
__device__ __forceinline__ int acgt_index(uint32_t c) {          // A C G T(U) -> 0 1 2 3, anything else -1
    if (c - 'a' < 26u) c -= 32;
    return c == 'A' ? 0 : c == 'C' ? 1 : c == 'G' ? 2 : (c == 'T' || c == 'U') ? 3 : -1;
}
__global__ void op_stats_kernel(const NsPieceMeta* pieces, const NsReadMeta* reads, const uint32_t* ops, uint32_t n,
                                DevRef ref, const uint8_t* seq, unsigned long long* st) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const NsPieceMeta pm = pieces[i];
    const NsReadMeta rm = reads[pm.read_slot];
    // pieces a read no longer owns (its piece list was replaced by ns_reemit) are not counted
    if (i < rm.piece_first || i >= rm.piece_first + rm.n_pieces) return;
    unsigned long long* ev = st + 8;
    unsigned long long* evlen = st + 16;
    unsigned long long* run_h = evlen + 3 * (NS_STATS_EV_CAP + 1);
    unsigned long long* first_h = run_h + (NS_STATS_RUN_CAP + 1);
    if (NS_PIECE_KIND(pm.kind) != NS_PIECE_SEGMENT) {
        atomicAdd(&st[4], 1ull);
        atomicAdd(&st[5], (unsigned long long)pm.out_len);
        return;
    }
    atomicAdd(&st[0], 1ull);
    atomicAdd(&st[1], (unsigned long long)pm.ref_len);
    uint64_t run = 0, ht = 0, n_ev = 0;
    bool first = true;
    // substituted / inserted bases are read back from the sequence: only when the event script is the emitted script
    // (-hp rewrites it) and the piece reads the reference forwards
    const bool bases_ok = seq != nullptr && pm.ev_off == pm.op_off && !(pm.kind & NS_PIECE_REF_REV);
    const uint64_t cstart = ref.chrom_off[pm.chrom];
    const uint32_t clen = (uint32_t)(ref.chrom_off[pm.chrom + 1] - cstart);
    const uint8_t* rs = seq ? seq + rm.seq_off : nullptr;
    uint32_t sub[16], insb[4];
#pragma unroll
    for (int k = 0; k < 16; ++k) sub[k] = 0;
#pragma unroll
    for (int k = 0; k < 4; ++k) insb[k] = 0;
    auto read_base = [&](uint32_t x) -> int {                    // base x of the read in the reference's orientation
        if (!rm.reversed) return acgt_index(rs[x]);
        const int b = acgt_index(rs[rm.seq_len - 1u - x]);
        return b < 0 ? b : 3 - b;
    };
    uint32_t o = pm.out_rel, rf = 0;
    for (uint32_t k = 0; k < pm.ev_n_ops; ++k) {         // the event script (== op script unless -hp rewrote it)
        uint32_t op = ops[pm.ev_off + k];
        uint32_t ty = op >> 28, len = op_len(op);
        if (ty == NS_OP_HT) {
            ht += len;
        } else if (ty == NS_OP_COPY) {
            run += len;
        } else if (ty <= NS_OP_DEL) {
            uint32_t t = ty - 1;   // 0 mis 1 ins 2 del
            atomicAdd(&ev[t], 1ull);
            atomicAdd(&ev[3 + t], (unsigned long long)len);
            atomicAdd(&evlen[t * (NS_STATS_EV_CAP + 1) + (len < NS_STATS_EV_CAP ? len : NS_STATS_EV_CAP)], 1ull);
            unsigned long long* h = first ? first_h : run_h;
            atomicAdd(&h[run < NS_STATS_RUN_CAP ? run : NS_STATS_RUN_CAP], 1ull);
            first = false;
            run = 0;
            ++n_ev;
            if (bases_ok && ty == NS_OP_MIS && len == 1) {
                uint32_t ab = pm.pos + rf;
                if (ab >= clen) ab -= clen;
                const int a = acgt_index(ref.bases[cstart + ab]), b = read_base(o);
                if (a >= 0 && b >= 0) ++sub[a * 4 + b];
            } else if (bases_ok && ty == NS_OP_INS) {
                for (uint32_t t2 = 0; t2 < len; ++t2) {
                    const int b = read_base(o + t2);
                    if (b >= 0) ++insb[b];
                }
            }
        }
        if (ty != NS_OP_DEL) o += len;
        if (ty == NS_OP_COPY || ty == NS_OP_MIS || ty == NS_OP_DEL) rf += len;
    }
    atomicAdd(&st[2], (unsigned long long)(pm.out_len - ht));
    atomicAdd(&st[3], (unsigned long long)ht);
    atomicAdd(&st[6], (unsigned long long)n_ev);
    if (n_ev) atomicAdd(&st[NS_STATS_EPR_OFF + (n_ev < NS_STATS_EPR_CAP ? n_ev : NS_STATS_EPR_CAP)], 1ull);
#pragma unroll
    for (int k = 0; k < 16; ++k)
        if (sub[k]) atomicAdd(&st[NS_STATS_SUB_OFF + k], (unsigned long long)sub[k]);
#pragma unroll
    for (int k = 0; k < 4; ++k)
        if (insb[k]) atomicAdd(&st[NS_STATS_INS_OFF + k], (unsigned long long)insb[k]);
}
// base composition of the batch's reads (as emitted; U counts as T): one warp per read
__global__ void base_comp_kernel(const NsReadMeta* reads, uint32_t n, const uint8_t* seq, unsigned long long* st) {
    const uint32_t w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (w >= n) return;
    const NsReadMeta rm = reads[w];
    const uint8_t* rs = seq + rm.seq_off;
    uint32_t c[4] = {0, 0, 0, 0};
    for (uint32_t x = lane; x < rm.seq_len; x += 32) {
        const int b = acgt_index(rs[x]);
        if (b >= 0) ++c[b];
    }
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        uint32_t v = c[k];
        for (int d = 16; d > 0; d >>= 1) v += __shfl_down_sync(0xffffffffu, v, d);
        if (lane == 0 && v) atomicAdd(&st[NS_STATS_COMP_OFF + k], (unsigned long long)v);
    }
}

cudaError_t upload(DevBuf& b, const void* src, size_t bytes, cudaStream_t s) {
    cudaError_t e = b.ensure(bytes ? bytes : 16);
    if (e != cudaSuccess) return e;
    if (bytes == 0) return cudaSuccess;
    return cudaMemcpyAsync(b.p, src, bytes, cudaMemcpyDefault, s);
}

// Base-quality sampler tables (emit_kernel.cuh:qual_pick).  The truncated log-normal of a state, floored to integers
// (model_base_qualities.py:9-20, 120-130), arrives as a 32-bit cdf; it is rounded to a 24-bit pmf (masses sum to 2^24) and
// laid out as a Walker alias table with QLUT_SIZE = 2048 slots of capacity 2^13: slot j holds a primary and an alias
// quality and the primary's share t of the slot, so a 24-bit draw (11 bits slot, 13 bits u) returns the primary iff u < t.
// All arithmetic is in integers: the table realises the 24-bit pmf exactly.  Entry: t << 19 | alias char << 8 | primary
// char (ASCII = q + 33); a slot that holds one quality only stores it twice (t is then irrelevant).
void build_qlut(const uint32_t cdf32[NS_N_QUAL_STATES][NS_QUAL_SLOTS], std::vector<uint32_t>& lut) {
    lut.assign((size_t)NS_N_QUAL_STATES * QLUT_SIZE, 0);
    const uint32_t cap = 1u << (24 - QLUT_BITS);
    for (int s = 0; s < NS_N_QUAL_STATES; ++s) {
        uint32_t c24[NS_QUAL_SLOTS];
        for (int q = 0; q < NS_QUAL_SLOTS; ++q) {
            uint64_t v = ((uint64_t)cdf32[s][q] + 128u) >> 8;
            c24[q] = (uint32_t)std::min<uint64_t>(v, 1u << 24);
            if (q && c24[q] < c24[q - 1]) c24[q] = c24[q - 1];
        }
        c24[NS_QUAL_SLOTS - 1] = 1u << 24;
        std::vector<uint32_t> mass(QLUT_SIZE, 0u), thr(QLUT_SIZE, 0u), prim(QLUT_SIZE), ali(QLUT_SIZE);
        for (int q = 0; q < NS_QUAL_SLOTS; ++q) mass[q] = c24[q] - (q ? c24[q - 1] : 0u);
        std::vector<uint32_t> small, large;
        for (uint32_t j = 0; j < QLUT_SIZE; ++j) {
            prim[j] = ali[j] = j;
            (mass[j] < cap ? small : large).push_back(j);
        }
        while (!small.empty() && !large.empty()) {
            const uint32_t a = small.back(), g = large.back();
            small.pop_back();
            large.pop_back();
            thr[a] = mass[a];                   // the rest of slot a, cap - mass[a], is taken from g
            ali[a] = g;
            mass[g] -= cap - mass[a];
            (mass[g] < cap ? small : large).push_back(g);
        }
        // what is left holds exactly `cap` (the masses are integers that sum to QLUT_SIZE * cap): pure slots
        for (uint32_t j : small) { thr[j] = 0; ali[j] = prim[j]; }
        for (uint32_t j : large) { thr[j] = 0; ali[j] = prim[j]; }
        // a slot index >= 94 is not a quality: such slots had no mass and are aliased entirely (thr 0)
        for (uint32_t j = 0; j < QLUT_SIZE; ++j) {
            uint32_t pq = prim[j], aq = ali[j];
            if (pq >= (uint32_t)NS_QUAL_SLOTS) pq = aq;
            if (aq >= (uint32_t)NS_QUAL_SLOTS) aq = pq;           // cannot happen: only slots with mass are aliases
            lut[(size_t)s * QLUT_SIZE + j] = (thr[j] << 19) | ((aq + 33u) << 8) | (pq + 33u);
        }
    }
}

}  // namespace

extern "C" {

int ns_create(int device, uint64_t seed, NsContext** out) {
    if (!out) return NS_EINVAL;
    *out = nullptr;
    NsContext* ctx = new NsContext();
    ctx->device = device;
    ctx->seed = seed;
    cudaError_t e = cudaSetDevice(device);
    if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking);
    for (int i = 0; i < 6 && e == cudaSuccess; ++i) e = cudaEventCreate(&ctx->ev[i]);
    if (e == cudaSuccess) e = cudaHostAlloc((void**)&ctx->h_totals, 16 * sizeof(uint64_t), cudaHostAllocMapped);
    if (e == cudaSuccess) e = cudaHostGetDevicePointer((void**)&ctx->h_totals_dev, ctx->h_totals, 0);
    if (e == cudaSuccess && device >= 0 && device < 64 && !g_base[device]) {
        e = cudaEventCreate(&g_base[device]);
        if (e == cudaSuccess) e = cudaEventRecord(g_base[device], ctx->stream);
        if (e == cudaSuccess) e = cudaEventSynchronize(g_base[device]);
    }
    if (e == cudaSuccess) {
        cudaDeviceProp prop;
        e = cudaGetDeviceProperties(&prop, device);
        if (e == cudaSuccess) ctx->sm_count = prop.multiProcessorCount;
    }
    if (e != cudaSuccess) {
        fprintf(stderr, "nanosim_b200: ns_create failed: %s\n", cudaGetErrorString(e));
        delete ctx;
        return NS_ECUDA;
    }
    *out = ctx;
    return NS_OK;
}

int ns_destroy(NsContext* ctx) {
    if (!ctx) return NS_EINVAL;
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    DevBuf* bufs[] = {&ctx->ref_bases, &ctx->ref_off, &ctx->ref_packed, &ctx->ref_pk_off, &ctx->ref_exc, &ctx->alias, &ctx->qlut, &ctx->reads, &ctx->pieces,
                      &ctx->ops, &ctx->seq, &ctx->qual, &ctx->nseg, &ctx->npieces, &ctx->piece_first, &ctx->scan_in,
                      &ctx->scan_out, &ctx->scan_tmp, &ctx->counter, &ctx->totals, &ctx->stats, &ctx->sort_keys, &ctx->sort_vals,
                      &ctx->sort_tmp, &ctx->hp_off, &ctx->ref_species, &ctx->ref_circular, &ctx->ref_sp_off, &ctx->sp_bases_dev, &ctx->kde2d_x, &ctx->kde2d_y,
                      &ctx->expr_alias, &ctx->expr_chrom, &ctx->chrom_polya, &ctx->split_base, &ctx->split_extra, &ctx->split_ckpt, &ctx->hp_keys};
    if (ctx->borrowed) {            // shared with the parent: drop the pointers without freeing
        DevBuf* shared[] = {&ctx->ref_bases, &ctx->ref_off, &ctx->ref_packed, &ctx->ref_pk_off, &ctx->ref_exc, &ctx->alias, &ctx->qlut, &ctx->ref_species,
                            &ctx->ref_circular, &ctx->ref_sp_off, &ctx->kde2d_x, &ctx->kde2d_y, &ctx->expr_alias,
                            &ctx->expr_chrom, &ctx->chrom_polya};
        for (DevBuf* b : shared) { b->p = nullptr; b->cap = 0; }
        for (auto& k : ctx->kde) { k.p = nullptr; k.cap = 0; }
    }
    if (ctx->trx_sorted_owned) ctx->trx_sorted.release();
    for (DevBuf* b : bufs) b->release();
    for (auto& k : ctx->kde) k.release();
    if (ctx->h_totals) cudaFreeHost(ctx->h_totals);
    if (ctx->pack_host) cudaFreeHost(ctx->pack_host);
    ctx->pack_dev.release();
    if (ctx->ev_pack) cudaEventDestroy(ctx->ev_pack);
    if (ctx->ev_block) cudaEventDestroy(ctx->ev_block);
    for (auto& e : ctx->ev)
        if (e) cudaEventDestroy(e);
    if (ctx->stream) cudaStreamDestroy(ctx->stream);
    delete ctx;
    return NS_OK;
}

const char* ns_last_error(const NsContext* ctx) { return ctx ? ctx->err.c_str() : "null context"; }

int ns_clone(NsContext* parent, NsContext** out) {
    if (!parent || !out) return NS_EINVAL;
    NsContext* c = nullptr;
    int rc = ns_create(parent->device, parent->seed, &c);
    if (rc != NS_OK) return rc;
    c->borrowed = true;
    c->have_ref = parent->have_ref;
    c->have_model = parent->have_model;
    c->have_cfg = parent->have_cfg;
    c->ref_bases = parent->ref_bases;      // plain pointer copies; ns_destroy() of a clone does not free them
    c->ref_off = parent->ref_off;
    c->ref_packed = parent->ref_packed;
    c->ref_pk_off = parent->ref_pk_off;
    c->ref_exc = parent->ref_exc;
    c->trx_sorted = parent->trx_sorted;        // shared; trx_sorted_owned stays false
    c->alias = parent->alias;
    c->qlut = parent->qlut;
    c->ref_species = parent->ref_species;
    c->ref_circular = parent->ref_circular;
    c->ref_sp_off = parent->ref_sp_off;
    c->kde2d_x = parent->kde2d_x;
    c->kde2d_y = parent->kde2d_y;
    c->expr_alias = parent->expr_alias;
    c->expr_chrom = parent->expr_chrom;
    c->chrom_polya = parent->chrom_polya;
    c->have_expr = parent->have_expr;
    c->h_sp_off = parent->h_sp_off;
    c->abun = parent->abun;
    c->abun_inflated = parent->abun_inflated;
    c->species_bases.assign(parent->abun.size(), 0.0);      // every clone is its own worker (own running totals)
    for (int i = 0; i < 5; ++i) c->kde[i] = parent->kde[i];
    c->dref = parent->dref;
    c->h_chrom_off = parent->h_chrom_off;
    c->dmodel = parent->dmodel;
    c->hmodel = parent->hmodel;
    c->hcfg = parent->hcfg;
    c->dcfg = parent->dcfg;
    *out = c;
    return NS_OK;
}

int ns_set_reference(NsContext* ctx, const NsReference* ref) {
    if (!ctx || !ref || !ref->bases || !ref->chrom_off || ref->n_chrom == 0)
        return fail(ctx, NS_EINVAL, "ns_set_reference: null argument or empty reference");
    if (ctx->borrowed) return fail(ctx, NS_ESTATE, "ns_set_reference: a cloned context shares its parent's reference");
    CK(cudaSetDevice(ctx->device));
    ctx->h_chrom_off.resize(ref->n_chrom + 1);
    CK(cudaMemcpy(ctx->h_chrom_off.data(), ref->chrom_off, (ref->n_chrom + 1) * sizeof(uint64_t), cudaMemcpyDefault));
    if (ctx->h_chrom_off[0] != 0 || ctx->h_chrom_off[ref->n_chrom] != ref->n_bases)
        return fail(ctx, NS_EINVAL, "ns_set_reference: chrom_off must start at 0 and end at n_bases");
    for (uint32_t i = 0; i < ref->n_chrom; ++i) {
        uint64_t len = ctx->h_chrom_off[i + 1] - ctx->h_chrom_off[i];
        if (ctx->h_chrom_off[i + 1] < ctx->h_chrom_off[i] || len > 0xffffffffull)
            return fail(ctx, NS_EINVAL, "ns_set_reference: chromosome %u has an invalid length", i);
    }
    CK(upload(ctx->ref_bases, ref->bases, ref->n_bases, ctx->stream));
    CK(upload(ctx->ref_off, ctx->h_chrom_off.data(), (ref->n_chrom + 1) * sizeof(uint64_t), ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    ctx->dref.bases = ctx->ref_bases.as<uint8_t>();
    ctx->dref.chrom_off = ctx->ref_off.as<uint64_t>();
    {   // 2-bit copy + exception counts for the emit kernel's fast path
        std::vector<uint64_t> pk(ref->n_chrom + 1);
        uint64_t words = 1;                                             // guard word in front
        for (uint32_t i = 0; i < ref->n_chrom; ++i) {
            pk[i] = words;
            words += (ctx->h_chrom_off[i + 1] - ctx->h_chrom_off[i] + 15) / 16;
        }
        pk[ref->n_chrom] = words;
        const uint64_t n_blocks = ((words + 2) >> REF_EXC_BLOCK_SHIFT) + 2;
        CK(ctx->ref_packed.ensure((size_t)(words + 2 + 64) * 4));      // + slack: a 64-word window copy may start at the last word
        CK(ctx->ref_exc.ensure((size_t)(2 * n_blocks + 2) * 4 + 16));
        CK(upload(ctx->ref_pk_off, pk.data(), pk.size() * sizeof(uint64_t), ctx->stream));
        CK(cudaMemsetAsync(ctx->ref_packed.p, 0, (size_t)(words + 2 + 64) * 4, ctx->stream));
        CK(cudaMemsetAsync(ctx->ref_exc.p, 0, (size_t)(2 * n_blocks + 2) * 4 + 16, ctx->stream));
        uint32_t* cnt = ctx->ref_exc.as<uint32_t>() + n_blocks + 1;     // counts behind the prefix array
        unsigned long long* other = (unsigned long long*)(ctx->ref_exc.as<uint32_t>() + ((2 * n_blocks + 2 + 1) & ~1ull));
        pack_reference_kernel<<<(unsigned)((words + 255) / 256), 256, 0, ctx->stream>>>(
            ctx->ref_bases.as<uint8_t>(), ctx->ref_off.as<uint64_t>(), ctx->ref_pk_off.as<uint64_t>(), ref->n_chrom,
            ctx->ref_packed.as<uint32_t>(), words, cnt, other);
        CK(cudaGetLastError());
        size_t tmp = 0;
        CK(cub::DeviceScan::ExclusiveSum(nullptr, tmp, cnt, ctx->ref_exc.as<uint32_t>(), (int)(n_blocks + 1), ctx->stream));
        CK(ctx->scan_tmp.ensure(tmp));
        CK(cub::DeviceScan::ExclusiveSum(ctx->scan_tmp.p, tmp, cnt, ctx->ref_exc.as<uint32_t>(), (int)(n_blocks + 1), ctx->stream));
        unsigned long long h_other = 0;
        CK(cudaMemcpyAsync(&h_other, other, sizeof h_other, cudaMemcpyDeviceToHost, ctx->stream));
        CK(cudaStreamSynchronize(ctx->stream));
        ctx->dref.packed = ctx->ref_packed.as<uint32_t>();
        ctx->dref.pk_words = words + 2;
        ctx->dref.pk_off = ctx->ref_pk_off.as<uint64_t>();
        ctx->dref.exc_pre = ctx->ref_exc.as<uint32_t>();
        ctx->dref.all_iupac = h_other == 0 ? 1u : 0u;
    }
    ctx->dref.genome_len = ref->n_bases;
    ctx->dref.n_chrom = ref->n_chrom;
    ctx->dref.n_species = 0;
    ctx->dref.chrom_species = nullptr;
    ctx->dref.chrom_circular = nullptr;
    ctx->dref.species_chrom_off = nullptr;
    ctx->dref.expr_alias = nullptr;
    ctx->dref.expr_chrom = nullptr;
    ctx->dref.n_expressed = 0;
    ctx->dref.chrom_has_polya = nullptr;
    ctx->dref.trx_len_sorted = nullptr;
    ctx->dref.trx_len_idx = nullptr;
    ctx->dref.n_trx_sorted = 0;
    ctx->have_expr = false;
    if (ref->n_species > 0) {
        if (!ref->chrom_species || !ref->chrom_circular)
            return fail(ctx, NS_EINVAL, "ns_set_reference: n_species > 0 needs chrom_species and chrom_circular");
        std::vector<uint32_t> sp(ref->n_chrom);
        std::vector<uint8_t> circ(ref->n_chrom);
        CK(cudaMemcpy(sp.data(), ref->chrom_species, sp.size() * 4, cudaMemcpyDefault));
        CK(cudaMemcpy(circ.data(), ref->chrom_circular, circ.size(), cudaMemcpyDefault));
        ctx->h_sp_off.assign(ref->n_species + 1, 0);
        for (uint32_t i = 0; i < ref->n_chrom; ++i) {
            if (sp[i] >= ref->n_species || (i > 0 && sp[i] < sp[i - 1]))
                return fail(ctx, NS_EINVAL, "ns_set_reference: chromosomes must be grouped by species in species order");
            ctx->h_sp_off[sp[i] + 1] = i + 1;
        }
        for (uint32_t k = 1; k <= ref->n_species; ++k) {
            if (ctx->h_sp_off[k] == 0) ctx->h_sp_off[k] = ctx->h_sp_off[k - 1];
            if (ctx->h_sp_off[k] == ctx->h_sp_off[k - 1]) return fail(ctx, NS_EINVAL, "ns_set_reference: species %u has no chromosome", k - 1);
        }
        CK(upload(ctx->ref_species, sp.data(), sp.size() * 4, ctx->stream));
        CK(upload(ctx->ref_circular, circ.data(), circ.size(), ctx->stream));
        CK(upload(ctx->ref_sp_off, ctx->h_sp_off.data(), ctx->h_sp_off.size() * 4, ctx->stream));
        CK(cudaStreamSynchronize(ctx->stream));
        ctx->dref.n_species = ref->n_species;
        ctx->dref.chrom_species = ctx->ref_species.as<uint32_t>();
        ctx->dref.chrom_circular = ctx->ref_circular.as<uint8_t>();
        ctx->dref.species_chrom_off = ctx->ref_sp_off.as<uint32_t>();
    }
    ctx->have_ref = true;
    ctx->have_batch = false;
    ctx->opt_ok[0] = ctx->opt_ok[1] = false;
    return NS_OK;
}

int ns_set_model(NsContext* ctx, const NsModel* m) {
    if (!ctx || !m) return fail(ctx, NS_EINVAL, "ns_set_model: null argument");
    if (ctx->borrowed) return fail(ctx, NS_ESTATE, "ns_set_model: a cloned context shares its parent's model");
    if (m->n_match_bins == 0 || m->n_match_bins > NS_MAX_BINS || m->n_tables != 4 + m->n_match_bins)
        return fail(ctx, NS_EINVAL, "ns_set_model: need 1..%d match bins and 4+bins alias tables", NS_MAX_BINS);
    if (!m->alias_prob || !m->alias_idx || !m->alias_desc || !m->match_bin_lo || !m->match_bin_hi)
        return fail(ctx, NS_EINVAL, "ns_set_model: null table pointer");
    CK(cudaSetDevice(ctx->device));
    ctx->hmodel = *m;
    DevModel& d = ctx->dmodel;
    const NsKde* src[5] = {&m->kde_aligned, &m->kde_ht, &m->kde_ht_ratio, &m->kde_unaligned, &m->kde_gap};
    DevKde* dst[5] = {&d.aligned, &d.ht, &d.ratio, &d.unaligned, &d.gap};
    for (int i = 0; i < 5; ++i) {
        if (src[i]->n && !src[i]->data) return fail(ctx, NS_EINVAL, "ns_set_model: KDE %d has n>0 but no data", i);
        CK(upload(ctx->kde[i], src[i]->data, (size_t)src[i]->n * sizeof(float), ctx->stream));
        dst[i]->data = ctx->kde[i].as<float>();
        dst[i]->n = src[i]->n;
        dst[i]->bw = src[i]->bandwidth;
    }
    if ((m->kde_aligned.n == 0 && m->n_kde2d == 0) || m->kde_ht.n == 0 || m->kde_ht_ratio.n == 0)
        return fail(ctx, NS_EINVAL, "ns_set_model: aligned (or 2-D aligned) / ht / ht_ratio KDEs are required");
    d.kde2d_x = d.kde2d_y = nullptr;
    d.n_kde2d = 0;
    d.kde2d_bw = 0.f;
    if (m->n_kde2d > 0) {
        if (!m->kde2d_x || !m->kde2d_y) return fail(ctx, NS_EINVAL, "ns_set_model: n_kde2d > 0 needs kde2d_x / kde2d_y");
        std::vector<float> hx(m->n_kde2d);
        CK(cudaMemcpy(hx.data(), m->kde2d_x, hx.size() * sizeof(float), cudaMemcpyDefault));
        for (uint32_t i = 1; i < m->n_kde2d; ++i)
            if (hx[i] < hx[i - 1]) return fail(ctx, NS_EINVAL, "ns_set_model: kde2d_x must be sorted ascending");
        CK(upload(ctx->kde2d_x, m->kde2d_x, (size_t)m->n_kde2d * sizeof(float), ctx->stream));
        CK(upload(ctx->kde2d_y, m->kde2d_y, (size_t)m->n_kde2d * sizeof(float), ctx->stream));
        d.kde2d_x = ctx->kde2d_x.as<float>();
        d.kde2d_y = ctx->kde2d_y.as<float>();
        d.n_kde2d = m->n_kde2d;
        d.kde2d_bw = m->kde2d_bandwidth;
    }
    // interleave (prob, alias) so that one 8-byte load serves a draw
    std::vector<uint32_t> hp(m->alias_len), hi(m->alias_len), desc(2 * m->n_tables);
    CK(cudaMemcpy(hp.data(), m->alias_prob, hp.size() * 4, cudaMemcpyDefault));
    CK(cudaMemcpy(hi.data(), m->alias_idx, hi.size() * 4, cudaMemcpyDefault));
    CK(cudaMemcpy(desc.data(), m->alias_desc, desc.size() * 4, cudaMemcpyDefault));
    std::vector<uint2> inter(m->alias_len);
    for (uint32_t i = 0; i < m->alias_len; ++i) inter[i] = make_uint2(hp[i], hi[i]);
    CK(upload(ctx->alias, inter.data(), inter.size() * sizeof(uint2), ctx->stream));
    d.alias = ctx->alias.as<uint2>();
    for (uint32_t t = 0; t < m->n_tables; ++t) {
        d.tab_off[t] = desc[2 * t];
        d.tab_n[t] = desc[2 * t + 1];
        if (d.tab_n[t] == 0 || (uint64_t)d.tab_off[t] + d.tab_n[t] > m->alias_len)
            return fail(ctx, NS_EINVAL, "ns_set_model: alias table %u out of range", t);
    }
    std::vector<uint32_t> blo(m->n_match_bins), bhi(m->n_match_bins);
    CK(cudaMemcpy(blo.data(), m->match_bin_lo, blo.size() * 4, cudaMemcpyDefault));
    CK(cudaMemcpy(bhi.data(), m->match_bin_hi, bhi.size() * 4, cudaMemcpyDefault));
    d.n_bins = m->n_match_bins;
    for (uint32_t b = 0; b < d.n_bins; ++b) {
        d.bin_lo[b] = blo[b];
        d.bin_hi[b] = bhi[b];
    }
    memcpy(d.trans, m->trans, sizeof d.trans);
    d.strandness = m->strandness_rate;
    d.seg_p = m->segment_mean > 1.0f ? 1.0 / (double)m->segment_mean : 1.0;
    std::vector<uint32_t> lut;
    build_qlut(m->qual_cdf, lut);
    CK(upload(ctx->qlut, lut.data(), lut.size() * 4, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    ctx->have_model = true;
    ctx->have_batch = false;
    ctx->opt_ok[0] = ctx->opt_ok[1] = false;
    return NS_OK;
}

int ns_configure(NsContext* ctx, const NsRunConfig* cfg) {
    if (!ctx || !cfg) return fail(ctx, NS_EINVAL, "ns_configure: null argument");
    if (cfg->mode > 2) return fail(ctx, NS_EINVAL, "ns_configure: mode must be 0 (genome), 1 (metagenome) or 2 (transcriptome)");
    if (cfg->mode == 2 && cfg->chimeric) return fail(ctx, NS_EINVAL, "ns_configure: transcriptome reads are not chimeric");
    if (cfg->mode == 2 && ctx->have_model && ctx->dmodel.n_kde2d == 0)
        return fail(ctx, NS_ESTATE, "ns_configure: transcriptome mode needs _aligned_region_2d.pkl in the model");
    if (cfg->mode == 1 && ctx->have_ref && ctx->dref.n_species == 0)
        return fail(ctx, NS_ESTATE, "ns_configure: metagenome mode needs a reference with species information");
    if (cfg->mode == 1 && cfg->kmer_bias != 0) return fail(ctx, NS_EINVAL, "ns_configure: -hp/-k is not offered in metagenome mode");
    if (cfg->max_len < cfg->min_len) return fail(ctx, NS_EINVAL, "Maximum read length must be longer than Minimum read length!");
    if (cfg->perfect && cfg->chimeric) return fail(ctx, NS_EINVAL, "Perfect reads cannot be chimeric");
    if ((cfg->median_len != 0.0) != (cfg->sd_len != 0.0))
        return fail(ctx, NS_EINVAL, "Please provide both mean and standard deviation of read length!");
    if (cfg->median_len != 0.0 && cfg->chimeric) return fail(ctx, NS_EINVAL, "Lognormal distributed reads cannot be chimeric!");
    if (cfg->median_len < 0.0 || cfg->sd_len < 0.0) return fail(ctx, NS_EINVAL, "ns_configure: negative -med/-sd");
    if (cfg->kmer_bias != 0 && ctx->have_model && !ctx->hmodel.has_hp)
        return fail(ctx, NS_ESTATE, "ns_configure: -hp/-k needs _hp_lengths_model_parameters.tsv in the model");
    if (cfg->kmer_bias == 1) return fail(ctx, NS_EINVAL, "ns_configure: -k must be >= 2 (every base is a run of length 1)");
    ctx->hcfg = *cfg;
    ctx->dcfg.circular = cfg->circular;
    ctx->dcfg.perfect = cfg->perfect;
    ctx->dcfg.fastq = cfg->fastq;
    ctx->dcfg.chimeric = cfg->chimeric;
    ctx->dcfg.kmer_bias = cfg->kmer_bias;
    ctx->dcfg.metagenome = cfg->mode == 1 ? 1u : 0u;
    ctx->dcfg.transcriptome = cfg->mode == 2 ? 1u : 0u;
    ctx->dcfg.uracil = (cfg->flags & NS_FLAG_URACIL) ? 1u : 0u;
    ctx->dcfg.kde2d_n = cfg->kde2d_sample ? cfg->kde2d_sample : 1u;
    ctx->dcfg.trx_records = cfg->trx_records;
    if (cfg->mode == 2 && ctx->have_ref) {
        // records sorted by length for the unaligned reads' transcript draw (plan_kernel.cuh:draw_position_trx)
        const uint32_t nrec = cfg->trx_records ? std::min(cfg->trx_records, ctx->dref.n_chrom) : ctx->dref.n_chrom;
        if (ctx->dref.n_trx_sorted != nrec || !ctx->dref.trx_len_sorted) {
            std::vector<uint32_t> idx(nrec), both(2 * (size_t)nrec);
            for (uint32_t i = 0; i < nrec; ++i) idx[i] = i;
            const std::vector<uint64_t>& off = ctx->h_chrom_off;
            std::stable_sort(idx.begin(), idx.end(), [&](uint32_t x, uint32_t y) { return off[x + 1] - off[x] < off[y + 1] - off[y]; });
            for (uint32_t i = 0; i < nrec; ++i) {
                both[i] = (uint32_t)(off[idx[i] + 1] - off[idx[i]]);
                both[nrec + i] = idx[i];
            }
            CK(cudaSetDevice(ctx->device));
            if (!ctx->trx_sorted_owned) {          // a clone must not grow (= free) its parent's buffer
                ctx->trx_sorted.p = nullptr;
                ctx->trx_sorted.cap = 0;
            }
            CK(upload(ctx->trx_sorted, both.data(), both.size() * 4, ctx->stream));
            CK(cudaStreamSynchronize(ctx->stream));
            ctx->trx_sorted_owned = true;
            ctx->dref.trx_len_sorted = ctx->trx_sorted.as<uint32_t>();
            ctx->dref.trx_len_idx = ctx->trx_sorted.as<uint32_t>() + nrec;
            ctx->dref.n_trx_sorted = nrec;
        }
    }
    ctx->dcfg.polya_scale = cfg->polya_scale;
    ctx->dcfg.min_len = cfg->min_len;
    ctx->dcfg.max_len = cfg->max_len;
    ctx->dcfg.seed = ctx->seed;
    ctx->dcfg.median_len = cfg->median_len;
    ctx->dcfg.sd_len = cfg->sd_len;
    ctx->have_cfg = true;
    ctx->have_batch = false;
    ctx->opt_ok[0] = ctx->opt_ok[1] = false;
    return NS_OK;
}

int ns_set_abundance(NsContext* ctx, const double* abun, const double* abun_inflated, uint32_t n_species) {
    if (!ctx || !abun || n_species == 0) return fail(ctx, NS_EINVAL, "ns_set_abundance: null argument");
    if (!ctx->have_ref || ctx->dref.n_species != n_species)
        return fail(ctx, NS_ESTATE, "ns_set_abundance: the reference has %u species, got %u", ctx ? ctx->dref.n_species : 0, n_species);
    ctx->abun.assign(abun, abun + n_species);
    if (abun_inflated) ctx->abun_inflated.assign(abun_inflated, abun_inflated + n_species);
    else ctx->abun_inflated.assign(n_species, 0.0);
    ctx->species_bases.assign(n_species, 0.0);
    return NS_OK;
}

int ns_set_expression(NsContext* ctx, const NsExpression* ex) {
    if (!ctx || !ex || !ex->alias_prob || !ex->alias_idx || !ex->expr_chrom || ex->n_expressed == 0)
        return fail(ctx, NS_EINVAL, "ns_set_expression: null argument or no expressed transcript");
    if (!ctx->have_ref) return fail(ctx, NS_ESTATE, "ns_set_expression: set the reference transcriptome first");
    if (ctx->borrowed) return fail(ctx, NS_ESTATE, "ns_set_expression: a cloned context shares its parent's tables");
    CK(cudaSetDevice(ctx->device));
    const uint32_t n = ex->n_expressed;
    std::vector<uint32_t> hp(n), hi(n), hc(n);
    CK(cudaMemcpy(hp.data(), ex->alias_prob, n * 4, cudaMemcpyDefault));
    CK(cudaMemcpy(hi.data(), ex->alias_idx, n * 4, cudaMemcpyDefault));
    CK(cudaMemcpy(hc.data(), ex->expr_chrom, n * 4, cudaMemcpyDefault));
    std::vector<uint2> inter(n);
    for (uint32_t i = 0; i < n; ++i) {
        if (hi[i] >= n || hc[i] >= ctx->dref.n_chrom) return fail(ctx, NS_EINVAL, "ns_set_expression: index out of range at %u", i);
        inter[i] = make_uint2(hp[i], hi[i]);
    }
    CK(upload(ctx->expr_alias, inter.data(), inter.size() * sizeof(uint2), ctx->stream));
    CK(upload(ctx->expr_chrom, hc.data(), hc.size() * 4, ctx->stream));
    ctx->dref.expr_alias = ctx->expr_alias.as<uint2>();
    ctx->dref.expr_chrom = ctx->expr_chrom.as<uint32_t>();
    ctx->dref.n_expressed = n;
    ctx->dref.chrom_has_polya = nullptr;
    if (ex->chrom_has_polya) {
        CK(upload(ctx->chrom_polya, ex->chrom_has_polya, ctx->dref.n_chrom, ctx->stream));
        ctx->dref.chrom_has_polya = ctx->chrom_polya.as<uint8_t>();
    }
    CK(cudaStreamSynchronize(ctx->stream));
    ctx->have_expr = true;
    ctx->have_batch = false;
    return NS_OK;
}

// ---- assign_species (:758-811) on the host: a sequential greedy pass over the batch's segments.
namespace {
struct HostRng {       // splitmix64 stream keyed by (seed, batch id); only drives the species choices
    uint64_t s;
    uint64_t next() {
        uint64_t z = (s += 0x9E3779B97F4A7C15ull);
        z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
        z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
        return z ^ (z >> 31);
    }
    double uniform() { return (double)(next() >> 11) * (1.0 / 9007199254740992.0); }
    uint32_t below(uint32_t n) { return (uint32_t)(((unsigned __int128)next() * n) >> 64); }
};

// reads: order = reads with the most segments first (stable), then single-segment reads by decreasing length;
// quota[s] = total_bases * abun[s] / sum(abun) - current[s]; every segment takes a uniformly random species whose quota
// still fits it (else any species with quota left); later segments of a chimeric read stay in the previous species with
// probability abun_inflated[prev] % (:793-797).
// seg_len[j]: drawn length of segment j (segments of read i are first[i] .. first[i] + n_seg[i]); by_len: read slots by
// decreasing total drawn length (the device's stable radix sort, = decreasing segment length for single-segment reads).
// The species whose quota exceeds a length are a PREFIX of the species sorted by quota, so a pick is a binary search + one
// uniform draw, and charging a quota moves one species a few places down that order.
void assign_species_host(const std::vector<uint32_t>& n_seg, const std::vector<uint32_t>& first, const std::vector<uint32_t>& seg_len,
                         const std::vector<uint32_t>& by_len, std::vector<uint32_t>& seg_species, const std::vector<double>& abun,
                         const std::vector<double>& inflated, const std::vector<double>& current, HostRng& rng) {
    const uint32_t n = (uint32_t)n_seg.size(), S = (uint32_t)abun.size();
    std::vector<uint32_t> order;
    order.reserve(n);
    for (uint32_t k = NS_MAX_SEGMENTS; k >= 2; --k)                        // most segments first, stable
        for (uint32_t i = 0; i < n; ++i)
            if (n_seg[i] == k) order.push_back(i);
    for (uint32_t q = 0; q < n; ++q)
        if (n_seg[by_len[q]] <= 1) order.push_back(by_len[q]);
    double to_add = 0, cur = 0, tot_abun = 0;
    for (uint32_t v : seg_len) to_add += v;
    for (uint32_t s = 0; s < S; ++s) {
        cur += current[s];
        tot_abun += abun[s];
    }
    std::vector<double> quota(S);
    for (uint32_t s = 0; s < S; ++s) quota[s] = (to_add + cur) * abun[s] / tot_abun - current[s];
    std::vector<uint32_t> ord(S), where(S);                                // species by decreasing quota, and its inverse
    for (uint32_t s = 0; s < S; ++s) ord[s] = s;
    std::stable_sort(ord.begin(), ord.end(), [&](uint32_t x, uint32_t y) { return quota[x] > quota[y]; });
    for (uint32_t k = 0; k < S; ++k) where[ord[k]] = k;
    auto count_above = [&](double v) {                                      // species with quota > v
        uint32_t lo = 0, hi = S;
        while (lo < hi) {
            const uint32_t mid = (lo + hi) >> 1;
            if (quota[ord[mid]] > v) lo = mid + 1; else hi = mid;
        }
        return lo;
    };
    auto pick_above = [&](double v, int exclude) -> uint32_t {              // uniform among them, `exclude` left out
        const uint32_t c = count_above(v);
        const bool ex_in = exclude >= 0 && where[exclude] < c;
        const uint32_t m = c - (ex_in ? 1u : 0u);
        if (m == 0) return 0xffffffffu;
        uint32_t r = rng.below(m);
        if (ex_in && r >= where[exclude]) ++r;
        return ord[r];
    };
    auto pick = [&](double len, int exclude) -> uint32_t {
        uint32_t sp = pick_above(len, exclude);                             // quota - len > 0
        if (sp == 0xffffffffu && exclude < 0) sp = pick_above(0.0, -1);     // else any species with quota left
        return sp;
    };
    auto charge = [&](uint32_t sp, double len) {
        quota[sp] -= len;
        uint32_t k = where[sp];
        while (k + 1 < S && quota[ord[k + 1]] > quota[sp]) {               // keeps `ord` sorted
            ord[k] = ord[k + 1];
            where[ord[k]] = k;
            ++k;
        }
        ord[k] = sp;
        where[sp] = k;
    };
    for (uint32_t oi = 0; oi < n; ++oi) {
        const uint32_t i = order[oi];
        int pre = -1;
        for (uint32_t q = 0; q < n_seg[i]; ++q) {
            const double len = seg_len[first[i] + q];
            uint32_t sp;
            if (q == 0) {
                sp = pick(len, -1);
            } else {
                const double p = rng.uniform() * 100.0;
                uint32_t other = pick(len, pre);
                if (p <= inflated[pre] && quota[pre] > 0) sp = (uint32_t)pre;
                else if (p > inflated[pre] && other != 0xffffffffu) sp = other;
                else sp = pick(len, -1);
            }
            if (sp == 0xffffffffu) sp = ord[0];           // cannot happen while sum(quota) >= len; keep it total anyway
            seg_species[first[i] + q] = sp;
            charge(sp, len);
            pre = (int)sp;
        }
    }
}

// segment lengths down / species up: 4 bytes per segment instead of whole NsPieceMeta records
__global__ void gather_segment_req(const NsPieceMeta* pieces, const uint32_t* piece_first, const uint32_t* n_seg, uint32_t n_reads,
                                   const uint32_t* seg_first, uint32_t* out) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_reads) return;
    const uint32_t pf = piece_first ? piece_first[i] : i, ns = n_seg ? n_seg[i] : 1u, sf = seg_first ? seg_first[i] : i;
    for (uint32_t q = 0; q < ns; ++q) out[sf + q] = pieces[pf + 2 * q].ref_req;
}
__global__ void scatter_segment_species(NsPieceMeta* pieces, const uint32_t* piece_first, const uint32_t* n_seg, uint32_t n_reads,
                                        const uint32_t* seg_first, const uint32_t* species) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_reads) return;
    const uint32_t pf = piece_first ? piece_first[i] : i, ns = n_seg ? n_seg[i] : 1u, sf = seg_first ? seg_first[i] : i;
    for (uint32_t q = 0; q < ns; ++q) pieces[pf + 2 * q].chrom = species[sf + q];    // species id travels in `chrom` until the position is drawn
}

__global__ void species_bases_kernel(const NsPieceMeta* pieces, uint32_t n, const uint32_t* chrom_species, double* acc) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n && NS_PIECE_KIND(pieces[i].kind) == NS_PIECE_SEGMENT) atomicAdd(&acc[chrom_species[pieces[i].chrom]], (double)pieces[i].ref_len);
}
}  // namespace

static int exclusive_scan_u64(NsContext* ctx, const uint64_t* in, uint64_t* out, uint32_t n) {
    size_t tmp = 0;
    CK(cub::DeviceScan::ExclusiveSum(nullptr, tmp, in, out, (int)n, ctx->stream));
    CK(ctx->scan_tmp.ensure(tmp));
    CK(cub::DeviceScan::ExclusiveSum(ctx->scan_tmp.p, tmp, in, out, (int)n, ctx->stream));
    return NS_OK;
}

namespace {
// Launch geometry: NANOSIM_B200_EMIT_BLOCKS_PER_SM / NANOSIM_B200_PLAN_BLOCKS_PER_SM cap the persistent grids (0 / unset =
// as many blocks as fit), so that kernels of overlapped contexts can share an SM instead of queueing.
// emit_kernel over `n_pieces` pieces of the context's current batch (all of them, or the ones `order` lists)
int launch_emit(NsContext* ctx, int kind, uint64_t first_read_id, uint32_t n_pieces, const uint32_t* order,
                const uint32_t* abort_flag = nullptr, bool split = true, cudaEvent_t emit_begin = nullptr) {
    cudaStream_t st = ctx->stream;
    EmitArgs ea;
    ea.ref = ctx->dref;
    ea.cfg = ctx->dcfg;
    ea.kind = (uint32_t)kind;
    ea.first_id = first_read_id;
    ea.reads = ctx->reads.as<NsReadMeta>();
    ea.pieces = ctx->pieces.as<NsPieceMeta>();
    ea.ops = ctx->ops.as<uint32_t>();
    ea.n_pieces = n_pieces;
    ea.seq = ctx->seq.as<uint8_t>();
    ea.qual = ctx->qual.as<uint8_t>();
    ea.qlut = ctx->qlut.as<uint32_t>();
    ea.force_exact = (ctx->hcfg.flags & NS_FLAG_EMIT_EXACT) ? 1u : 0u;
    ea.abort = abort_flag;
    ea.counter = ctx->counter.as<uint32_t>();
    ea.order = order;
    CK(cudaMemsetAsync(ctx->counter.p, 0, 64, st));
    {   // long pieces become several work items.  Capacity: every extra item stands for EMIT_SPLIT - 16 or more output bytes
        // of its piece, and a batch's output fits the sequence buffer (checked on the device for sync-free batches).
        static const int no_split = env_int("NANOSIM_B200_NO_SPLIT", 0);
        const uint32_t cap_extra = (uint32_t)std::min<size_t>(ctx->seq.cap / (EMIT_SPLIT - 16u) + 64u, 0x7fffffffu);
        ea.split_base = nullptr;
        ea.extra = nullptr;
        ea.ckpt = nullptr;
        ea.n_extra = ctx->counter.as<uint32_t>() + 4;
        ea.cap_extra = cap_extra;
        if (split && !no_split && n_pieces && !(ctx->hcfg.flags & NS_FLAG_EMIT_WHOLE)) {     // (`order` lists pieces below n_pieces whenever split is asked for)
            CK(ctx->split_base.ensure((size_t)n_pieces * sizeof(uint32_t)));
            CK(ctx->split_extra.ensure((size_t)cap_extra * sizeof(uint2)));
            CK(ctx->split_ckpt.ensure((size_t)cap_extra * sizeof(uint4)));
            CK(cudaMemsetAsync(ctx->split_extra.p, 0xff, (size_t)cap_extra * sizeof(uint2), st));
            SplitArgs sa;
            sa.reads = ea.reads;
            sa.pieces = ea.pieces;
            sa.ops = ea.ops;
            sa.n_pieces = n_pieces;
            sa.split_base = ctx->split_base.as<uint32_t>();
            sa.extra = ctx->split_extra.as<uint2>();
            sa.ckpt = ctx->split_ckpt.as<uint4>();
            sa.n_extra = ctx->counter.as<uint32_t>() + 4;
            sa.cap_extra = cap_extra;
            sa.abort = abort_flag;
            const unsigned sblocks = std::min<unsigned>((n_pieces + 7u) / 8u, (unsigned)ctx->sm_count * 8u);
            split_kernel<<<sblocks, 256, 0, st>>>(sa);
            if (emit_begin) CK(cudaEventRecord(emit_begin, st));     // the emit phase of the batch timings starts after the split
            ea.split_base = sa.split_base;
            ea.extra = sa.extra;
            ea.ckpt = sa.ckpt;
        }
    }
    const size_t ring_bytes = (size_t)EMIT_WARPS * EMIT_RING * sizeof(uint4) + 512 + EMIT_WINDOW_SMEM;
    if (ctx->hcfg.fastq) {
        size_t smem = ring_bytes + (size_t)NS_N_QUAL_STATES * QLUT_SIZE * 4;
        CK(cudaFuncSetAttribute(emit_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        int per_sm = 1;
        CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, emit_kernel<true>, EMIT_WARPS * 32, smem));
        static const int cap_sm = env_int("NANOSIM_B200_EMIT_BLOCKS_PER_SM", 0);
        if (cap_sm > 0) per_sm = std::min(per_sm, cap_sm);
        unsigned blocks = std::min<unsigned>((n_pieces + EMIT_WARPS - 1) / EMIT_WARPS, (unsigned)(ctx->sm_count * std::max(per_sm, 1)));
        emit_kernel<true><<<blocks, EMIT_WARPS * 32, smem, st>>>(ea);
    } else {
        size_t smem = ring_bytes;
        int per_sm = 1;
        CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, emit_kernel<false>, EMIT_WARPS * 32, smem));
        static const int cap_sm = env_int("NANOSIM_B200_EMIT_BLOCKS_PER_SM", 0);
        if (cap_sm > 0) per_sm = std::min(per_sm, cap_sm);
        unsigned blocks = std::min<unsigned>((n_pieces + EMIT_WARPS - 1) / EMIT_WARPS, (unsigned)(ctx->sm_count * std::max(per_sm, 1)));
        emit_kernel<false><<<blocks, EMIT_WARPS * 32, smem, st>>>(ea);
    }
    CK(cudaGetLastError());
    return NS_OK;
}

__global__ void hp_piece_keys(const NsPieceMeta* pieces, uint32_t n, uint32_t* keys, uint32_t* vals) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    keys[i] = NS_PIECE_KIND(pieces[i].kind) == NS_PIECE_SEGMENT ? pieces[i].ref_len : 0u;
    vals[i] = i;
}

__global__ void replace_reads(NsReadMeta* reads, const uint32_t* slots, const NsReadMeta* repl, uint32_t n, uint32_t n_reads) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n && slots[i] < n_reads) {
        NsReadMeta r = repl[i];
        r.seq_off = reads[slots[i]].seq_off;         // the bytes are overwritten in place
        reads[slots[i]] = r;
    }
}
__global__ void iota_from(uint32_t* v, uint32_t n, uint32_t first) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) v[i] = first + i;
}
}  // namespace

int ns_simulate(NsContext* ctx, int kind, uint64_t first_read_id, uint32_t n_reads, NsBatchInfo* info) {
    if (!ctx) return NS_EINVAL;
    if (!ctx->have_ref || !ctx->have_model || !ctx->have_cfg)
        return fail(ctx, NS_ESTATE, "ns_simulate: reference, model and run configuration must be set first");
    if (kind != NS_KIND_ALIGNED && kind != NS_KIND_UNALIGNED) return fail(ctx, NS_EINVAL, "ns_simulate: bad kind %d", kind);
    if (kind == NS_KIND_UNALIGNED && ctx->dmodel.unaligned.n == 0 && ctx->hcfg.median_len == 0.0)
        return fail(ctx, NS_ESTATE, "ns_simulate: model has no unaligned-length KDE");
    if (kind == NS_KIND_ALIGNED && ctx->hcfg.chimeric && ctx->dmodel.gap.n == 0)
        return fail(ctx, NS_ESTATE, "ns_simulate: chimeric simulation needs the gap-length KDE");
    if (ctx->dcfg.transcriptome && kind == NS_KIND_ALIGNED && !ctx->have_expr)
        return fail(ctx, NS_ESTATE, "ns_simulate: transcriptome mode needs ns_set_expression");
    if (ctx->hcfg.fastq && !ctx->hmodel.has_qual)
        return fail(ctx, NS_ESTATE, "ns_simulate: --fastq needs base-quality parameters in the model");
    if (ctx->hcfg.max_len > 0x0fffffffu) ctx->dcfg.max_len = 0x0fffffffu;
    CK(cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->stream;
    ctx->have_batch = false;
    memset(&ctx->last, 0, sizeof ctx->last);
    if (n_reads == 0) {
        if (info) *info = ctx->last;
        ctx->have_batch = true;
        ctx->last_kind = kind;
        return NS_OK;
    }
    const uint32_t n = n_reads;
    uint32_t launches = 0;
    const bool fast_unaligned = kind == NS_KIND_UNALIGNED && !(ctx->hcfg.flags & NS_FLAG_UNALIGNED_SCRIPTS);
    const unsigned tb = 256, gb = (n + tb - 1) / tb;
    CK(ctx->reads.ensure((size_t)n * sizeof(NsReadMeta)));
    CK(ctx->counter.ensure(64));
    CK(ctx->totals.ensure(16 * sizeof(uint64_t)));
    CK(ctx->scan_in.ensure((size_t)n * (2 * NS_MAX_SEGMENTS) * sizeof(uint64_t)));
    CK(ctx->scan_out.ensure((size_t)n * (2 * NS_MAX_SEGMENTS) * sizeof(uint64_t)));
    CK(cudaMemsetAsync(ctx->totals.p, 0, 16 * sizeof(uint64_t), st));
    CK(cudaEventRecord(ctx->ev[0], st));

    // ---- pieces per read
    const bool chim = (kind == NS_KIND_ALIGNED) && ctx->hcfg.chimeric;
    // one submission, no host round trip: same kind of batch as one already sized, no chimeric piece count, species
    // assignment, homopolymer pass or scripted unaligned reads in the way
    static const bool no_opt = getenv("NANOSIM_B200_SYNC_BATCHES") != nullptr;
    const bool optimistic = !no_opt && ctx->opt_ok[kind] && n <= ctx->opt_n[kind] && !chim && !(ctx->dcfg.metagenome && kind == NS_KIND_ALIGNED) &&
                            !(ctx->hcfg.kmer_bias > 0 && kind == NS_KIND_ALIGNED) && (kind == NS_KIND_ALIGNED || fast_unaligned) &&
                            ctx->ops.cap > 64 && ctx->seq.cap > 64 && (!ctx->hcfg.fastq || ctx->qual.cap >= ctx->seq.cap);
    const uint64_t ops_cap = ctx->ops.cap / sizeof(uint32_t), seq_cap = ctx->seq.cap;
    uint64_t* d_totals = ctx->totals.as<uint64_t>();
    const uint32_t* d_abort = optimistic ? (const uint32_t*)(d_totals + NS_T_ABORT) : nullptr;
    uint32_t n_pieces = n;
    const uint32_t* d_nseg = nullptr;
    const uint32_t* d_pfirst = nullptr;
    if (chim) {
        CK(ctx->nseg.ensure((size_t)n * 4));
        CK(ctx->npieces.ensure((size_t)n * 4));
        CK(ctx->piece_first.ensure((size_t)n * 4));
        segments_kernel<<<gb, tb, 0, st>>>(ctx->dmodel, ctx->dcfg, (uint32_t)kind, first_read_id, n, ctx->nseg.as<uint32_t>(),
                                           ctx->npieces.as<uint32_t>());
        widen_u32<<<gb, tb, 0, st>>>(ctx->npieces.as<uint32_t>(), n, ctx->scan_in.as<uint64_t>());
        int rc = exclusive_scan_u64(ctx, ctx->scan_in.as<uint64_t>(), ctx->scan_out.as<uint64_t>(), n);
        if (rc) return rc;
        narrow_u64<<<gb, tb, 0, st>>>(ctx->scan_out.as<uint64_t>(), n, ctx->piece_first.as<uint32_t>());
        last_total<<<1, 32, 0, st>>>(ctx->scan_in.as<uint64_t>(), ctx->scan_out.as<uint64_t>(), n, ctx->totals.as<uint64_t>(), 0);
        launches += 6;
        publish_totals<<<1, 32, 0, st>>>(ctx->totals.as<uint64_t>(), ctx->h_totals_dev);
        CK(cudaStreamSynchronize(st));
        n_pieces = (uint32_t)ctx->h_totals[0];
        d_nseg = ctx->nseg.as<uint32_t>();
        d_pfirst = ctx->piece_first.as<uint32_t>();
    }
    CK(ctx->pieces.ensure((size_t)(n_pieces + 1) * sizeof(NsPieceMeta)));
    CK(cudaMemsetAsync(ctx->pieces.p, 0, (size_t)(n_pieces + 1) * sizeof(NsPieceMeta), st));
    const unsigned gp = (n_pieces + tb - 1) / tb;

    // ---- generation-0 lengths -> op-slot capacities (scan) and processing order (longest reads first)
    CK(ctx->sort_keys.ensure((size_t)n * 8));
    CK(ctx->sort_vals.ensure((size_t)n * 8));
    uint32_t* keys_in = ctx->sort_keys.as<uint32_t>();
    uint32_t* keys_out = keys_in + n;
    uint32_t* vals_in = ctx->sort_vals.as<uint32_t>();
    uint32_t* vals_out = vals_in + n;
    const bool exact_only = (kind == NS_KIND_UNALIGNED) && !fast_unaligned;
    lengths_kernel<<<gb, tb, 0, st>>>(ctx->dmodel, ctx->dcfg, (uint32_t)kind, first_read_id, n, d_nseg, d_pfirst,
                                      ctx->pieces.as<NsPieceMeta>(), 1.0f / std::max(1.0f, ctx->hmodel.mean_ref_per_event),
                                      std::min(8.0f, std::max(1.0f, ctx->hmodel.ref_per_event_cv)),
                                      exact_only ? 1u : 0u, ctx->scan_in.as<uint64_t>(), keys_in, vals_in);
    CK(cudaGetLastError());
    {
        int rc = exclusive_scan_u64(ctx, ctx->scan_in.as<uint64_t>(), ctx->scan_out.as<uint64_t>(), n_pieces);
        if (rc) return rc;
    }
    scatter_piece_off<<<gp, tb, 0, st>>>(ctx->pieces.as<NsPieceMeta>(), n_pieces, ctx->scan_out.as<uint64_t>());
    last_total<<<1, 32, 0, st>>>(ctx->scan_in.as<uint64_t>(), ctx->scan_out.as<uint64_t>(), n_pieces, ctx->totals.as<uint64_t>(), 4);
    set_sentinel_off<<<1, 32, 0, st>>>(ctx->pieces.as<NsPieceMeta>(), n_pieces, ctx->totals.as<uint64_t>() + 4);
    {
        size_t tmp = 0;
        CK(cub::DeviceRadixSort::SortPairsDescending(nullptr, tmp, keys_in, keys_out, vals_in, vals_out, (int)n, 0, 32, st));
        CK(ctx->sort_tmp.ensure(tmp));
        CK(cub::DeviceRadixSort::SortPairsDescending(ctx->sort_tmp.p, tmp, keys_in, keys_out, vals_in, vals_out, (int)n, 0, 32, st));
    }
    // primary script area = the capped slots (+ a bump pool for re-drawn unaligned reads, uread_kernel.cuh)
    uint64_t primary_ops = 0;
    capacity_stage_a<<<1, 32, 0, st>>>(d_totals, fast_unaligned ? 1u : 0u, optimistic ? ops_cap : ~0ull);
    if (!optimistic) {
        publish_totals<<<1, 32, 0, st>>>(d_totals, ctx->h_totals_dev);
        CK(cudaStreamSynchronize(st));
        primary_ops = ctx->h_totals[NS_T_PRIMARY];
        CK(ctx->ops.ensure((size_t)(primary_ops + 4) * sizeof(uint32_t)));
    }
    uint32_t batch_reversed = 0;
    if (ctx->dcfg.metagenome && kind == NS_KIND_ALIGNED) {
        // ---- assign_species (:758-811): sequential greedy quota fill over this batch's segments, on the host.  Down: the
        //      drawn segment lengths and the reads' order by length (4 B each); up: one species per segment.
        if (ctx->abun.size() != ctx->dref.n_species) return fail(ctx, NS_ESTATE, "ns_simulate: call ns_set_abundance first");
        std::vector<uint32_t> hseg(n, 1u), hfirst(n), by_len(n);
        uint32_t n_segs = n;
        if (chim) {
            CK(cudaMemcpyAsync(hseg.data(), ctx->nseg.p, (size_t)n * 4, cudaMemcpyDeviceToHost, st));
            CK(cudaStreamSynchronize(st));
            n_segs = 0;
            for (uint32_t i = 0; i < n; ++i) {
                hfirst[i] = n_segs;
                n_segs += hseg[i];
            }
        } else {
            for (uint32_t i = 0; i < n; ++i) hfirst[i] = i;
        }
        CK(ctx->hp_off.ensure((size_t)(n + n_segs) * 4));                 // scratch: segment starts + per-segment values
        uint32_t* d_sfirst = ctx->hp_off.as<uint32_t>();
        uint32_t* d_segval = d_sfirst + n;
        if (chim) CK(cudaMemcpyAsync(d_sfirst, hfirst.data(), (size_t)n * 4, cudaMemcpyHostToDevice, st));
        gather_segment_req<<<gb, tb, 0, st>>>(ctx->pieces.as<NsPieceMeta>(), d_pfirst, d_nseg, n, chim ? d_sfirst : nullptr, d_segval);
        std::vector<uint32_t> seg_len(n_segs), seg_species(n_segs, 0u);
        CK(cudaMemcpyAsync(seg_len.data(), d_segval, (size_t)n_segs * 4, cudaMemcpyDeviceToHost, st));
        CK(cudaMemcpyAsync(by_len.data(), vals_out, (size_t)n * 4, cudaMemcpyDeviceToHost, st));
        CK(cudaStreamSynchronize(st));
        HostRng hr{ctx->seed * 0x9E3779B97F4A7C15ull ^ (first_read_id + 0x1234567ull)};
        batch_reversed = hr.uniform() > (double)ctx->dmodel.strandness ? 1u : 0u;      // once per batch (:860)
        assign_species_host(hseg, hfirst, seg_len, by_len, seg_species, ctx->abun, ctx->abun_inflated, ctx->species_bases, hr);
        CK(cudaMemcpyAsync(d_segval, seg_species.data(), (size_t)n_segs * 4, cudaMemcpyHostToDevice, st));
        scatter_segment_species<<<gb, tb, 0, st>>>(ctx->pieces.as<NsPieceMeta>(), d_pfirst, d_nseg, n, chim ? d_sfirst : nullptr, d_segval);
        CK(cudaGetLastError());
        CK(cudaStreamSynchronize(st));                                    // seg_species must outlive the copy
    }
    CK(cudaEventRecord(ctx->ev[1], st));
    launches += 10;

    // ---- plan: rejection loops, positions, edit scripts (single pass)
    PlanArgs pa;
    pa.m = ctx->dmodel;
    pa.ref = ctx->dref;
    pa.cfg = ctx->dcfg;
    pa.kind = (uint32_t)kind;
    pa.first_id = first_read_id;
    pa.n_reads = n;
    pa.n_seg = d_nseg;
    pa.piece_first = d_pfirst;
    pa.reads = ctx->reads.as<NsReadMeta>();
    pa.pieces = ctx->pieces.as<NsPieceMeta>();
    pa.ops = ctx->ops.as<uint32_t>();
    pa.order = vals_out;
    pa.counter = ctx->counter.as<uint32_t>();
    pa.n_flagged = (uint32_t*)(ctx->totals.as<uint64_t>() + 5);
    pa.batch_reversed = batch_reversed;
    pa.abort = d_abort;
    const unsigned plan_tb = 128;
    // 2 resident blocks per SM: measured faster than 4 (the register limit) both alone (6.3-6.8 vs 8.2-8.6 ms on the config-2
    // batch) and next to another context's emit kernel, which then still finds room on every SM
    // (transcripts are short -- 1-2 kb reads, no long tail, two passes: there the register limit of 4 blocks is better)
    static const int plan_per_sm_env = env_int("NANOSIM_B200_PLAN_BLOCKS_PER_SM", 0);
    const int plan_per_sm = plan_per_sm_env > 0 ? plan_per_sm_env : (ctx->dcfg.transcriptome ? 4 : 2);
    unsigned plan_blocks = std::min<unsigned>((n + plan_tb - 1) / plan_tb, (unsigned)ctx->sm_count * (unsigned)std::max(1, plan_per_sm));
    // unaligned reads without NS_FLAG_UNALIGNED_SCRIPTS: warp-per-read evaluation (uread_kernel.cuh), same outputs
    UreadArgs ua;
    ua.m = ctx->dmodel;
    ua.ref = ctx->dref;
    ua.cfg = ctx->dcfg;
    ua.first_id = first_read_id;
    ua.n_reads = n;
    ua.reads = pa.reads;
    ua.pieces = pa.pieces;
    ua.ops = pa.ops;
    ua.order = vals_out;
    ua.counter = pa.counter;
    ua.n_flagged = pa.n_flagged;
    ua.pool_cursor = (unsigned long long*)(ctx->totals.as<uint64_t>() + 7);
    ua.pool = d_totals + NS_T_POOL;             // {base, size} of the bump pool, written by capacity_stage_a
    ua.abort = d_abort;
    static const int cta_min = env_int("NANOSIM_B200_UREAD_CTA_MIN", (int)UREAD_CTA_MIN_LEN);
    ua.cta_min_len = (ctx->hcfg.flags & NS_FLAG_EMIT_WHOLE) ? 0u : (uint32_t)std::max(cta_min, 0);
    const unsigned ublocks = std::min<unsigned>((n + UREAD_WARPS - 1) / UREAD_WARPS, (unsigned)ctx->sm_count * 8u);
    if (chim && !ctx->hcfg.perfect) {
        // chimeric gaps of every read's first attempt, a warp per read (uread_kernel.cuh:gap_kernel)
        GapArgs ga;
        ga.m = ctx->dmodel;
        ga.cfg = ctx->dcfg;
        ga.kind = (uint32_t)kind;
        ga.first_id = first_read_id;
        ga.n_reads = n;
        ga.n_seg = d_nseg;
        ga.piece_first = d_pfirst;
        ga.pieces = pa.pieces;
        ga.ops = pa.ops;
        ga.counter = pa.counter;
        ga.abort = d_abort;
        CK(cudaMemsetAsync(ctx->counter.p, 0, 64, st));
        gap_kernel<<<ublocks, UREAD_WARPS * 32, 0, st>>>(ga);
        CK(cudaGetLastError());
        launches += 1;
    }
    CK(cudaMemsetAsync(ctx->counter.p, 0, 64, st));
    if (fast_unaligned) uread_kernel<false><<<ublocks, UREAD_WARPS * 32, 0, st>>>(ua);
    else plan_kernel<false><<<plan_blocks, plan_tb, 0, st>>>(pa);
    CK(cudaGetLastError());
    CK(cudaEventRecord(ctx->ev[2], st));
    launches += 1;

    // ---- 16-byte aligned sequence slots per read (scan); exact offsets for scripts that overflowed their slot
    gather_read_bytes<<<gb, tb, 0, st>>>(pa.reads, n, ctx->scan_in.as<uint64_t>());
    {
        int rc = exclusive_scan_u64(ctx, ctx->scan_in.as<uint64_t>(), ctx->scan_out.as<uint64_t>(), n);
        if (rc) return rc;
    }
    scatter_read_off<<<gb, tb, 0, st>>>(pa.reads, n, ctx->scan_out.as<uint64_t>());
    last_total<<<1, 32, 0, st>>>(ctx->scan_in.as<uint64_t>(), ctx->scan_out.as<uint64_t>(), n, ctx->totals.as<uint64_t>(), 2);
    sum_bases<<<std::min<unsigned>(gb, 1024u), tb, 0, st>>>(pa.reads, n, (unsigned long long*)(ctx->totals.as<uint64_t>() + 3));
    gather_flagged_ops<<<gp, tb, 0, st>>>(pa.pieces, pa.reads, n_pieces, ctx->scan_in.as<uint64_t>());
    {
        int rc = exclusive_scan_u64(ctx, ctx->scan_in.as<uint64_t>(), ctx->scan_out.as<uint64_t>(), n_pieces);
        if (rc) return rc;
    }
    last_total<<<1, 32, 0, st>>>(ctx->scan_in.as<uint64_t>(), ctx->scan_out.as<uint64_t>(), n_pieces, ctx->totals.as<uint64_t>(), 1);
    uint64_t overflow_ops = 0, seq_bytes = 0, total_bases = 0, n_ops = 0;
    uint32_t n_flagged = 0;
    if (optimistic) {
        capacity_stage_b<<<1, 32, 0, st>>>(d_totals, ops_cap, seq_cap);
        n_flagged = 1;                     // unknown without a round trip: the (normally empty) replay is always submitted
    } else {
        publish_totals<<<1, 32, 0, st>>>(d_totals, ctx->h_totals_dev);
        CK(cudaStreamSynchronize(st));
        overflow_ops = ctx->h_totals[1], seq_bytes = ctx->h_totals[2], total_bases = ctx->h_totals[3];
        n_flagged = (uint32_t)(ctx->h_totals[5] & 0xffffffffull);
        n_ops = primary_ops + overflow_ops;
        CK(ctx->seq.ensure((size_t)seq_bytes + 16));
        if (ctx->hcfg.fastq) CK(ctx->qual.ensure((size_t)seq_bytes + 16));
    }
    CK(cudaEventRecord(ctx->ev[3], st));
    launches += 9;
    if (n_flagged > 0) {
        // ---- rare: replay the flagged reads and write their scripts behind the primary area
        if (!optimistic) CK(ctx->ops.ensure_keep((size_t)(n_ops + 4) * sizeof(uint32_t), (size_t)primary_ops * sizeof(uint32_t), st));
        pa.ops = ctx->ops.as<uint32_t>();
        if (optimistic) scatter_flagged_off_dev<<<gp, tb, 0, st>>>(pa.pieces, pa.reads, n_pieces, ctx->scan_out.as<uint64_t>(), d_totals + NS_T_PRIMARY);
        else scatter_flagged_off<<<gp, tb, 0, st>>>(pa.pieces, pa.reads, n_pieces, ctx->scan_out.as<uint64_t>(), primary_ops);
        CK(cudaMemsetAsync(ctx->counter.p, 0, 64, st));
        ua.ops = pa.ops;
        if (fast_unaligned) uread_kernel<true><<<ublocks, UREAD_WARPS * 32, 0, st>>>(ua);
        else plan_kernel<true><<<plan_blocks, plan_tb, 0, st>>>(pa);
        CK(cudaGetLastError());
        launches += 2;
    }
    copy_ev_fields<<<gp, tb, 0, st>>>(pa.pieces, n_pieces);
    uint64_t n_ops_total = n_ops, seq_bytes_final = seq_bytes, total_bases_final = total_bases;
    if (ctx->hcfg.kmer_bias > 0 && kind == NS_KIND_ALIGNED && !ctx->hcfg.perfect) {
        // ---- homopolymer pass (hp_kernel.cuh): count, re-scan lengths and script offsets, write
        if (!ctx->hmodel.has_hp) return fail(ctx, NS_ESTATE, "ns_simulate: -hp/-k needs homopolymer parameters in the model");
        HpArgs ha;
        ha.ref = ctx->dref;
        ha.cfg = ctx->dcfg;
        ha.first_id = first_read_id;
        ha.reads = pa.reads;
        ha.pieces = pa.pieces;
        ha.n_pieces = n_pieces;
        ha.ops = pa.ops;
        ha.out_n_ops = ctx->scan_in.as<uint64_t>();
        ha.out_off = nullptr;
        memcpy(ha.hp, ctx->hmodel.hp, sizeof ha.hp);
        ha.hp_mis_rate = ctx->hmodel.hp_mis_rate;
        ha.counter = ctx->counter.as<uint32_t>();
        {   // segments longest first: the lanes of a warp walk segments of similar length, and the longest one starts first
            CK(ctx->hp_keys.ensure((size_t)n_pieces * 16));
            uint32_t* k_in = ctx->hp_keys.as<uint32_t>();
            uint32_t *k_out = k_in + n_pieces, *v_in = k_out + n_pieces, *v_out = v_in + n_pieces;
            hp_piece_keys<<<gp, tb, 0, st>>>(pa.pieces, n_pieces, k_in, v_in);
            size_t tmp = 0;
            CK(cub::DeviceRadixSort::SortPairsDescending(nullptr, tmp, k_in, k_out, v_in, v_out, (int)n_pieces, 0, 32, st));
            CK(ctx->sort_tmp.ensure(tmp));
            CK(cub::DeviceRadixSort::SortPairsDescending(ctx->sort_tmp.p, tmp, k_in, k_out, v_in, v_out, (int)n_pieces, 0, 32, st));
            ha.order = v_out;
        }
        ha.force_exact = (ctx->hcfg.flags & NS_FLAG_EMIT_EXACT) ? 1u : 0u;
        const unsigned hp_blocks = std::min<unsigned>((n_pieces + 127) / 128, (unsigned)ctx->sm_count * 16u);
        CK(cudaMemsetAsync(ctx->counter.p, 0, 64, st));
        hp_kernel<false><<<hp_blocks, 128, 0, st>>>(ha);
        CK(cudaGetLastError());
        hp_fix_reads<<<gb, tb, 0, st>>>(pa.reads, pa.pieces, n);
        CK(ctx->hp_off.ensure((size_t)n_pieces * sizeof(uint64_t)));
        {
            int rc = exclusive_scan_u64(ctx, ctx->scan_in.as<uint64_t>(), ctx->hp_off.as<uint64_t>(), n_pieces);
            if (rc) return rc;
        }
        last_total<<<1, 32, 0, st>>>(ctx->scan_in.as<uint64_t>(), ctx->hp_off.as<uint64_t>(), n_pieces, ctx->totals.as<uint64_t>(), 6);
        add_base_u64<<<gp, tb, 0, st>>>(ctx->hp_off.as<uint64_t>(), n_pieces, n_ops);
        CK(cudaMemsetAsync(ctx->totals.as<uint64_t>() + 3, 0, sizeof(uint64_t), st));
        gather_read_bytes<<<gb, tb, 0, st>>>(pa.reads, n, ctx->scan_in.as<uint64_t>());
        {
            int rc = exclusive_scan_u64(ctx, ctx->scan_in.as<uint64_t>(), ctx->scan_out.as<uint64_t>(), n);
            if (rc) return rc;
        }
        scatter_read_off<<<gb, tb, 0, st>>>(pa.reads, n, ctx->scan_out.as<uint64_t>());
        last_total<<<1, 32, 0, st>>>(ctx->scan_in.as<uint64_t>(), ctx->scan_out.as<uint64_t>(), n, ctx->totals.as<uint64_t>(), 2);
        sum_bases<<<std::min<unsigned>(gb, 1024u), tb, 0, st>>>(pa.reads, n, (unsigned long long*)(ctx->totals.as<uint64_t>() + 3));
        publish_totals<<<1, 32, 0, st>>>(ctx->totals.as<uint64_t>(), ctx->h_totals_dev);
        CK(cudaStreamSynchronize(st));
        n_ops_total = n_ops + ctx->h_totals[6];
        seq_bytes_final = ctx->h_totals[2];
        total_bases_final = ctx->h_totals[3];
        CK(ctx->ops.ensure_keep((size_t)(n_ops_total + 4) * sizeof(uint32_t), (size_t)n_ops * sizeof(uint32_t), st));
        CK(ctx->seq.ensure((size_t)seq_bytes_final + 16));
        if (ctx->hcfg.fastq) CK(ctx->qual.ensure((size_t)seq_bytes_final + 16));
        pa.ops = ctx->ops.as<uint32_t>();
        ha.ops = pa.ops;
        ha.out_off = ctx->hp_off.as<uint64_t>();
        CK(cudaMemsetAsync(ctx->counter.p, 0, 64, st));
        hp_kernel<true><<<hp_blocks, 128, 0, st>>>(ha);
        CK(cudaGetLastError());
        launches += 17;
    }
    CK(cudaEventRecord(ctx->ev[4], st));
    launches += 3;    // ev copy + split + emit

    // ---- emit
    {
        int rc = launch_emit(ctx, kind, first_read_id, n_pieces, chim ? nullptr : vals_out, d_abort, true, ctx->ev[4]);   // one piece per read: the plan's
        if (rc) return rc;                                                                       // longest-first order serves the emit too
    }
    CK(cudaGetLastError());
    CK(cudaEventRecord(ctx->ev[5], st));
    if (ctx->dcfg.metagenome && kind == NS_KIND_ALIGNED) {
        // current_species_bases[species] += len(new_seg) (:1004) for the next batch's quotas
        const uint32_t S = ctx->dref.n_species;
        CK(ctx->sp_bases_dev.ensure((size_t)S * sizeof(double)));
        CK(cudaMemsetAsync(ctx->sp_bases_dev.p, 0, (size_t)S * sizeof(double), st));
        species_bases_kernel<<<gp, tb, 0, st>>>(pa.pieces, n_pieces, ctx->dref.chrom_species, ctx->sp_bases_dev.as<double>());
        std::vector<double> add(S);
        CK(cudaMemcpyAsync(add.data(), ctx->sp_bases_dev.p, (size_t)S * sizeof(double), cudaMemcpyDeviceToHost, st));
        CK(cudaStreamSynchronize(st));
        for (uint32_t k = 0; k < S; ++k) ctx->species_bases[k] += add[k];
    }
    if (optimistic) publish_totals<<<1, 32, 0, st>>>(d_totals, ctx->h_totals_dev);
    CK(wait_stream(ctx, n >= 16384u));
    if (optimistic) {
        if (ctx->h_totals[NS_T_ABORT]) {
            // a buffer was too small for this batch: nothing was written past a capacity (the kernels saw the flag and
            // returned); run it again the sized way, which also re-establishes the capacities
            ctx->opt_ok[kind] = false;
            return ns_simulate(ctx, kind, first_read_id, n_reads, info);
        }
        n_ops_total = ctx->h_totals[NS_T_PRIMARY] + ctx->h_totals[1];
        seq_bytes_final = ctx->h_totals[2];
        total_bases_final = ctx->h_totals[3];
    } else if (!chim && !(ctx->dcfg.metagenome && kind == NS_KIND_ALIGNED)) {
        ctx->opt_ok[kind] = true;
        ctx->opt_n[kind] = n;
    }

    NsBatchInfo& bi = ctx->last;
    bi.seq_bytes = seq_bytes_final;
    bi.n_ops = n_ops_total;
    bi.total_bases = total_bases_final;
    bi.n_reads = n;
    bi.n_pieces = n_pieces;
    bi.n_launches = launches;
    float ms = 0;
    cudaEventElapsedTime(&ms, ctx->ev[0], ctx->ev[1]);
    bi.ms_setup = ms;
    cudaEventElapsedTime(&ms, ctx->ev[1], ctx->ev[2]);
    bi.ms_plan = ms;
    cudaEventElapsedTime(&ms, ctx->ev[2], ctx->ev[3]);
    bi.ms_scan = ms;
    cudaEventElapsedTime(&ms, ctx->ev[3], ctx->ev[4]);
    bi.ms_script = ms;
    cudaEventElapsedTime(&ms, ctx->ev[4], ctx->ev[5]);
    bi.ms_emit = ms;
    cudaEventElapsedTime(&ms, ctx->ev[0], ctx->ev[5]);
    bi.ms_total = ms;
    cudaEventElapsedTime(&ms, g_base[ctx->device], ctx->ev[0]);
    bi.t_begin_ms = ms;
    cudaEventElapsedTime(&ms, g_base[ctx->device], ctx->ev[5]);
    bi.t_end_ms = ms;
    ctx->last_kind = kind;
    ctx->last_first_id = first_read_id;
    ctx->have_batch = true;
    if (info) *info = bi;
    return NS_OK;
}

namespace {
// expands 2-bit bases (pack_bases_kernel) into ASCII with `nt` host threads
#if defined(__x86_64__)
// 32 characters from 8 packed bytes per step: every output byte gets its source byte (vpshufb), the three shifted copies
// bring the byte's other 2-bit fields down, constant masks keep field j & 3 at output byte j, a second vpshufb turns the
// indices into letters.  ~14 instructions per 32 bases instead of four table lookups: the expansion then runs at memory
// speed, which is what 8 GPU processes sharing one host need.
__attribute__((target("avx2"))) void unpack_range_avx2(const uint8_t* packed, uint8_t* seq, uint64_t lo, uint64_t hi, const char* abc) {
    const __m256i spread = _mm256_setr_epi8(0, 0, 0, 0, 1, 1, 1, 1, 2, 2, 2, 2, 3, 3, 3, 3, 4, 4, 4, 4, 5, 5, 5, 5, 6, 6, 6, 6, 7, 7, 7, 7);
    const __m256i m0 = _mm256_set1_epi32(0x00000003), m1 = _mm256_set1_epi32(0x00000300), m2 = _mm256_set1_epi32(0x00030000),
                  m3 = _mm256_set1_epi32(0x03000000);
    const __m256i letters = _mm256_setr_epi8(abc[0], abc[1], abc[2], abc[3], 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, abc[0], abc[1], abc[2], abc[3], 0,
                                             0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0);
    const bool aligned32 = (reinterpret_cast<uintptr_t>(seq) & 31u) == 0;
    for (uint64_t i = lo; i < hi; ++i) {              // unit: 32 characters
        const __m128i x = _mm_loadl_epi64(reinterpret_cast<const __m128i*>(packed + 8 * i));
        const __m256i src = _mm256_shuffle_epi8(_mm256_broadcastsi128_si256(x), spread);
        const __m256i idx = _mm256_or_si256(_mm256_or_si256(_mm256_and_si256(src, m0), _mm256_and_si256(_mm256_srli_epi16(src, 2), m1)),
                                            _mm256_or_si256(_mm256_and_si256(_mm256_srli_epi16(src, 4), m2), _mm256_and_si256(_mm256_srli_epi16(src, 6), m3)));
        const __m256i out = _mm256_shuffle_epi8(letters, idx);
        if (aligned32) _mm256_stream_si256(reinterpret_cast<__m256i*>(seq + 32 * i), out);     // written once, read much later
        else _mm256_storeu_si256(reinterpret_cast<__m256i*>(seq + 32 * i), out);
    }
    _mm_sfence();
}
#endif

void unpack_bases(const uint8_t* packed, uint8_t* seq, uint64_t seq_bytes, bool uracil, int nt) {
#if defined(__x86_64__)
    static const bool have_avx2 = __builtin_cpu_supports("avx2") && !getenv("NANOSIM_B200_NO_AVX2");
    if (have_avx2) {
        const char* abc2 = uracil ? "ACUG" : "ACTG";
        const uint64_t whole32 = seq_bytes / 32;
        nt = std::max(1, std::min(nt, 64));
        if (nt == 1 || whole32 < (1u << 14)) {
            unpack_range_avx2(packed, seq, 0, whole32, abc2);
        } else {
            std::vector<std::thread> th;
            for (int t = 0; t < nt; ++t) {
                const uint64_t lo = whole32 * t / nt, hi = whole32 * (t + 1) / nt;
                if (hi > lo) th.emplace_back(unpack_range_avx2, packed, seq, lo, hi, abc2);
            }
            for (auto& x : th) x.join();
        }
        for (uint64_t k = whole32 * 32; k < seq_bytes; ++k) seq[k] = (uint8_t)abc2[(packed[k >> 2] >> (2 * (k & 3))) & 3u];
        return;
    }
#endif
    // two packed bytes -> eight characters per table lookup (512 KB table per alphabet, built once)
    static std::vector<uint64_t> tables[2];
    static std::once_flag once[2];
    const char* abc = uracil ? "ACUG" : "ACTG";
    std::call_once(once[uracil ? 1 : 0], [&] {
        std::vector<uint64_t>& t = tables[uracil ? 1 : 0];
        t.resize(65536);
        for (uint32_t b = 0; b < 65536; ++b) {
            uint64_t w = 0;
            for (int j = 0; j < 8; ++j) w |= (uint64_t)(uint8_t)abc[(b >> (2 * j)) & 3u] << (8 * j);
            t[b] = w;
        }
    });
    const uint64_t* lut = tables[uracil ? 1 : 0].data();
    const uint64_t whole = seq_bytes / 8;             // 16-bit groups that expand to 8 in-range characters
    auto work = [&](uint64_t lo, uint64_t hi) {
        const bool aligned8 = (reinterpret_cast<uintptr_t>(seq) & 7u) == 0;
        for (uint64_t i = lo; i < hi; ++i) {
            uint16_t b;
            memcpy(&b, packed + 2 * i, 2);
            const uint64_t w = lut[b];
#if defined(__x86_64__)
            // streaming store: the destination is written once and read much later (no read-for-ownership traffic)
            if (aligned8) _mm_stream_si64(reinterpret_cast<long long*>(seq + 8 * i), (long long)w);
            else memcpy(seq + 8 * i, &w, 8);
#else
            memcpy(seq + 8 * i, &w, 8);
#endif
        }
#if defined(__x86_64__)
        _mm_sfence();
#endif
    };
    nt = std::max(1, std::min(nt, 64));
    if (nt == 1 || whole < (1u << 16)) {
        work(0, whole);
    } else {
        std::vector<std::thread> th;
        for (int t = 0; t < nt; ++t) {
            const uint64_t lo = whole * t / nt, hi = whole * (t + 1) / nt;
            if (hi > lo) th.emplace_back(work, lo, hi);
        }
        for (auto& x : th) x.join();
    }
    for (uint64_t k = whole * 8; k < seq_bytes; ++k) seq[k] = (uint8_t)abc[(packed[k >> 2] >> (2 * (k & 3))) & 3u];
}

// CPUs this process can really use: the affinity mask, capped by the container's cgroup CPU quota (cpu.max) -- the boxes of
// this pool show 128 logical CPUs and grant 16-24 cores of CPU time
unsigned effective_cpus() {
    unsigned n = std::thread::hardware_concurrency();
    cpu_set_t set;
    if (sched_getaffinity(0, sizeof set, &set) == 0) n = (unsigned)CPU_COUNT(&set);
    if (FILE* f = fopen("/sys/fs/cgroup/cpu.max", "r")) {
        char q[64] = {0};
        unsigned long long period = 0;
        if (fscanf(f, "%63s %llu", q, &period) == 2 && strcmp(q, "max") != 0 && period > 0) {
            const unsigned long long quota = strtoull(q, nullptr, 10);
            const unsigned cores = (unsigned)((quota + period / 2) / period);
            if (cores >= 1 && cores < n) n = cores;
        }
        fclose(f);
    }
    return n ? n : 4u;
}

int unpack_threads() {
    static const int n = [] {
        const char* e = getenv("NANOSIM_B200_UNPACK_THREADS");     // 0: copy the bases as ASCII (no packing)
        if (e && *e) return std::max(0, atoi(e));
        // packing only pays when the host can expand faster than PCIe delivers: one expanding thread per core this GPU
        // process can count on (torchrun exports LOCAL_WORLD_SIZE), at most 16; with fewer than 6 the bases travel as ASCII
        const char* lw = getenv("LOCAL_WORLD_SIZE");
        const unsigned ranks = (lw && *lw) ? (unsigned)std::max(1, atoi(lw)) : 1u;
        const unsigned per_rank = effective_cpus() / ranks;
        return per_rank >= 6u ? (int)std::min(per_rank, 16u) : 0;
    }();
    return n;
}
}  // namespace

int ns_unpack_bases(const uint8_t* packed, uint8_t* seq, uint64_t n_bases, int uracil, int threads) {
    if (!packed || !seq) return NS_EINVAL;
    unpack_bases(packed, seq, n_bases, uracil != 0, threads);
    return NS_OK;
}

int ns_fetch(NsContext* ctx, uint8_t* seq, uint8_t* qual, NsReadMeta* reads, NsPieceMeta* pieces, uint32_t* ops) {
    if (!ctx) return NS_EINVAL;
    if (!ctx->have_batch) return fail(ctx, NS_ESTATE, "ns_fetch: no simulated batch");
    CK(cudaSetDevice(ctx->device));
    const NsBatchInfo& bi = ctx->last;
    cudaStream_t st = ctx->stream;
    if (bi.n_reads == 0) return NS_OK;
    if (qual && !ctx->hcfg.fastq) return fail(ctx, NS_ESTATE, "ns_fetch: qualities requested but the run is not --fastq");
    // 2 bits per base only when reads cannot hold anything but A C G T/U: every reference byte is an IUPAC nucleotide code
    const int nt = ctx->dref.all_iupac ? unpack_threads() : 0;
    const bool packed = seq && nt > 0 && bi.seq_bytes >= (1u << 20);
    if (packed) {
        // bases: pack on the device, copy a quarter of the bytes, expand on the host while the other copies run
        const uint64_t n16 = (bi.seq_bytes + 15) / 16;            // the seq buffer has 16 bytes of slack
        CK(ctx->pack_dev.ensure((size_t)n16 * 4));
        if ((size_t)n16 * 4 > ctx->pack_host_cap) {
            if (ctx->pack_host) cudaFreeHost(ctx->pack_host);
            ctx->pack_host = nullptr;
            ctx->pack_host_cap = 0;
            const size_t want = (size_t)n16 * 4 + (size_t)n16 / 2 + 4096;
            CK(cudaHostAlloc((void**)&ctx->pack_host, want, cudaHostAllocDefault));
            ctx->pack_host_cap = want;
        }
        if (!ctx->ev_pack) CK(cudaEventCreateWithFlags(&ctx->ev_pack, cudaEventDisableTiming | (blocking_sync_wanted() ? (unsigned)cudaEventBlockingSync : 0u)));
        pack_bases_kernel<<<(unsigned)((n16 + 255) / 256), 256, 0, st>>>(ctx->seq.as<uint4>(), ctx->pack_dev.as<uint32_t>(), n16);
        CK(cudaGetLastError());
        CK(cudaMemcpyAsync(ctx->pack_host, ctx->pack_dev.p, (size_t)n16 * 4, cudaMemcpyDeviceToHost, st));
        CK(cudaEventRecord(ctx->ev_pack, st));
    } else if (seq) {
        CK(cudaMemcpyAsync(seq, ctx->seq.p, bi.seq_bytes, cudaMemcpyDeviceToHost, st));
    }
    if (qual) CK(cudaMemcpyAsync(qual, ctx->qual.p, bi.seq_bytes, cudaMemcpyDeviceToHost, st));
    if (reads) CK(cudaMemcpyAsync(reads, ctx->reads.p, (size_t)bi.n_reads * sizeof(NsReadMeta), cudaMemcpyDeviceToHost, st));
    if (pieces) CK(cudaMemcpyAsync(pieces, ctx->pieces.p, (size_t)bi.n_pieces * sizeof(NsPieceMeta), cudaMemcpyDeviceToHost, st));
    if (ops) CK(cudaMemcpyAsync(ops, ctx->ops.p, (size_t)bi.n_ops * sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
    static const bool trace = getenv("NANOSIM_B200_TRACE_FETCH") != nullptr;
    const auto t0 = std::chrono::steady_clock::now();
    auto ms_since = [&] { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count(); };
    double t_packed = 0, t_unpacked = 0;
    if (packed) {
        CK(cudaEventSynchronize(ctx->ev_pack));
        t_packed = ms_since();
        unpack_bases(ctx->pack_host, seq, bi.seq_bytes, ctx->dcfg.uracil != 0, nt);
        t_unpacked = ms_since();
    }
    CK(wait_stream(ctx, bi.seq_bytes >= (64u << 20)));
    if (trace)
        fprintf(stderr, "ns_fetch: %.2f GB bases; packed copy done after %.1f ms, expansion %.1f ms (%d threads), everything after %.1f ms\n",
                bi.seq_bytes / 1e9, t_packed, t_unpacked - t_packed, nt, ms_since());
    return NS_OK;
}

int ns_transfer_info(NsContext* ctx, uint32_t* packed_bases, uint32_t* n_threads) {
    if (!ctx) return NS_EINVAL;
    const int nt = (!ctx->have_ref || ctx->dref.all_iupac) ? unpack_threads() : 0;
    if (packed_bases) *packed_bases = nt > 0 ? 1u : 0u;
    if (n_threads) *n_threads = (uint32_t)nt;
    return NS_OK;
}

int ns_reemit(NsContext* ctx, const uint32_t* read_slots, const NsReadMeta* new_reads, uint32_t n_slots,
              const NsPieceMeta* new_pieces, uint32_t n_new_pieces, const uint32_t* new_ops, uint64_t n_new_ops) {
    if (!ctx) return NS_EINVAL;
    if (!ctx->have_batch || ctx->last_kind != NS_KIND_ALIGNED) return fail(ctx, NS_ESTATE, "ns_reemit: no aligned batch to patch");
    if (n_slots == 0 || n_new_pieces == 0) return NS_OK;
    if (!read_slots || !new_reads || !new_pieces || (n_new_ops && !new_ops)) return fail(ctx, NS_EINVAL, "ns_reemit: null argument");
    CK(cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->stream;
    NsBatchInfo& bi = ctx->last;
    const uint32_t old_np = bi.n_pieces;
    const uint64_t old_ops = bi.n_ops;
    // the host built absolute indices against these totals: check the ones that would corrupt memory
    for (uint32_t k = 0; k < n_slots; ++k) {
        if (read_slots[k] >= bi.n_reads) return fail(ctx, NS_EINVAL, "ns_reemit: read slot %u out of range", read_slots[k]);
        if (new_reads[k].piece_first < old_np || (uint64_t)new_reads[k].piece_first + new_reads[k].n_pieces > (uint64_t)old_np + n_new_pieces)
            return fail(ctx, NS_EINVAL, "ns_reemit: read %u does not point into the new pieces", read_slots[k]);
    }
    for (uint32_t k = 0; k < n_new_pieces; ++k) {
        const NsPieceMeta& p = new_pieces[k];
        if (p.read_slot >= bi.n_reads || p.chrom >= ctx->dref.n_chrom || p.op_off < old_ops || p.op_off + p.n_ops > old_ops + n_new_ops ||
            (uint64_t)p.pos + p.ref_len > ctx->h_chrom_off[p.chrom + 1] - ctx->h_chrom_off[p.chrom])
            return fail(ctx, NS_EINVAL, "ns_reemit: piece %u is inconsistent with the batch or the reference", k);
    }
    CK(ctx->pieces.ensure_keep((size_t)(old_np + n_new_pieces + 1) * sizeof(NsPieceMeta), (size_t)old_np * sizeof(NsPieceMeta), st));
    CK(ctx->ops.ensure_keep((size_t)(old_ops + n_new_ops + 4) * sizeof(uint32_t), (size_t)old_ops * sizeof(uint32_t), st));
    CK(cudaMemcpyAsync(ctx->pieces.as<NsPieceMeta>() + old_np, new_pieces, (size_t)n_new_pieces * sizeof(NsPieceMeta), cudaMemcpyHostToDevice, st));
    if (n_new_ops) CK(cudaMemcpyAsync(ctx->ops.as<uint32_t>() + old_ops, new_ops, (size_t)n_new_ops * sizeof(uint32_t), cudaMemcpyHostToDevice, st));
    // staging for the slots / replacement reads / piece order: the scan buffers are free between batches
    CK(ctx->scan_in.ensure((size_t)n_slots * sizeof(NsReadMeta)));
    CK(ctx->scan_out.ensure((size_t)n_slots * sizeof(uint32_t)));
    CK(ctx->sort_vals.ensure((size_t)n_new_pieces * sizeof(uint32_t)));
    CK(cudaMemcpyAsync(ctx->scan_in.p, new_reads, (size_t)n_slots * sizeof(NsReadMeta), cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(ctx->scan_out.p, read_slots, (size_t)n_slots * sizeof(uint32_t), cudaMemcpyHostToDevice, st));
    replace_reads<<<(n_slots + 255) / 256, 256, 0, st>>>(ctx->reads.as<NsReadMeta>(), ctx->scan_out.as<uint32_t>(),
                                                         ctx->scan_in.as<NsReadMeta>(), n_slots, bi.n_reads);
    iota_from<<<(n_new_pieces + 255) / 256, 256, 0, st>>>(ctx->sort_vals.as<uint32_t>(), n_new_pieces, old_np);
    CK(cudaGetLastError());
    {
        int rc = launch_emit(ctx, NS_KIND_ALIGNED, ctx->last_first_id, n_new_pieces, ctx->sort_vals.as<uint32_t>(), nullptr, false);
        if (rc) return rc;
    }
    CK(cudaStreamSynchronize(st));
    bi.n_pieces = old_np + n_new_pieces;
    bi.n_ops = old_ops + n_new_ops;
    bi.n_launches += 3;
    return NS_OK;
}

int ns_device_buffers(NsContext* ctx, const uint8_t** seq, const uint8_t** qual, const NsReadMeta** reads,
                      const NsPieceMeta** pieces, const uint32_t** ops) {
    if (!ctx) return NS_EINVAL;
    if (!ctx->have_batch) return fail(ctx, NS_ESTATE, "ns_device_buffers: no simulated batch");
    if (seq) *seq = ctx->seq.as<uint8_t>();
    if (qual) *qual = ctx->hcfg.fastq ? ctx->qual.as<uint8_t>() : nullptr;
    if (reads) *reads = ctx->reads.as<NsReadMeta>();
    if (pieces) *pieces = ctx->pieces.as<NsPieceMeta>();
    if (ops) *ops = ctx->ops.as<uint32_t>();
    return NS_OK;
}

int ns_op_stats(NsContext* ctx, uint64_t* out) {
    if (!ctx || !out) return NS_EINVAL;
    if (!ctx->have_batch) return fail(ctx, NS_ESTATE, "ns_op_stats: no simulated batch");
    CK(cudaSetDevice(ctx->device));
    const size_t bytes = (size_t)NS_STATS_WORDS * sizeof(uint64_t);
    CK(ctx->stats.ensure(bytes));
    CK(cudaMemsetAsync(ctx->stats.p, 0, bytes, ctx->stream));
    const uint32_t n = ctx->last.n_pieces;
    if (n) {
        op_stats_kernel<<<(n + 127) / 128, 128, 0, ctx->stream>>>(ctx->pieces.as<NsPieceMeta>(), ctx->reads.as<NsReadMeta>(),
                                                                   ctx->ops.as<uint32_t>(), n, ctx->dref, ctx->seq.as<uint8_t>(),
                                                                   (unsigned long long*)ctx->stats.p);
        const uint32_t nr = ctx->last.n_reads;
        if (nr && ctx->seq.p)
            base_comp_kernel<<<(nr + 3) / 4, 128, 0, ctx->stream>>>(ctx->reads.as<NsReadMeta>(), nr, ctx->seq.as<uint8_t>(),
                                                                     (unsigned long long*)ctx->stats.p);
        CK(cudaGetLastError());
    }
    CK(cudaMemcpyAsync(out, ctx->stats.p, bytes, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    return NS_OK;
}

// ---------------------------------------------------------------------------------------------------------
// host-side FASTA/FASTQ record formatting (simulator.py:1437-1443), multi-threaded memcpy-style assembly
// ---------------------------------------------------------------------------------------------------------
namespace {
// Sink of a formatter thread: either the caller's buffer, or a private chunk that is written with pwrite() at the right
// file position whenever it fills up (the records of one thread are contiguous in the output).
struct ChunkSink {
    char* out;                // buffer mode: start of the whole output
    int fd;                   // file mode: descriptor + offset of the output's first byte in the file
    uint64_t file_off;
    std::vector<char> buf;
    size_t used = 0;
    uint64_t start = 0;       // output position of buf[0]
    bool ok = true;
    ChunkSink(char* o, int f, uint64_t fo) : out(o), fd(f), file_off(fo) {
        if (fd >= 0) buf.resize(size_t(8) << 20);
    }
    char* reserve(uint64_t at, size_t n) {       // n bytes at output position `at` (positions only grow within a thread)
        if (fd < 0) return out + at;
        if (used + n > buf.size()) {
            flush();
            if (n > buf.size()) buf.resize(n);
        }
        if (used == 0) start = at;
        char* p = buf.data() + used;
        used += n;
        return p;
    }
    void flush() {
        size_t done = 0;
        while (fd >= 0 && done < used) {
            const ssize_t w = pwrite(fd, buf.data() + done, used - done, (off_t)(file_off + start + done));
            if (w <= 0) {
                ok = false;
                break;
            }
            done += (size_t)w;
        }
        used = 0;
    }
};

int64_t format_records_impl(const uint8_t* seq, const uint8_t* qual, const NsReadMeta* reads, uint32_t n_reads,
                            const char* names, const uint64_t* name_off, int fastq, char* out, uint64_t out_cap,
                            int n_threads, int fd, uint64_t file_off) {
    if (!seq || !reads || !names || !name_off || (fastq && !qual)) return NS_EINVAL;
    std::vector<uint64_t> off((size_t)n_reads + 1, 0);
    for (uint32_t i = 0; i < n_reads; ++i) {
        uint64_t nl = strlen(names + name_off[i]);
        uint64_t rec = 1 + nl + 1 + reads[i].seq_len + 1;
        if (fastq) rec += 2 + reads[i].seq_len + 1;
        off[i + 1] = off[i] + rec;
    }
    if (fd < 0) {
        if (!out) return (int64_t)off[n_reads];
        if (off[n_reads] > out_cap) return NS_ENOMEM;
    }
    int nt = std::max(1, std::min(n_threads, 64));
    std::vector<char> failed((size_t)nt, 0);
    auto work = [&](uint32_t lo, uint32_t hi, int tid) {
        ChunkSink sink(out, fd, file_off);
        for (uint32_t i = lo; i < hi; ++i) {
            char* p = sink.reserve(off[i], (size_t)(off[i + 1] - off[i]));
            const char* nm = names + name_off[i];
            size_t nl = strlen(nm);
            *p++ = fastq ? '@' : '>';
            memcpy(p, nm, nl);
            p += nl;
            *p++ = '\n';
            memcpy(p, seq + reads[i].seq_off, reads[i].seq_len);
            p += reads[i].seq_len;
            *p++ = '\n';
            if (fastq) {
                *p++ = '+';
                *p++ = '\n';
                memcpy(p, qual + reads[i].seq_off, reads[i].seq_len);
                p += reads[i].seq_len;
                *p++ = '\n';
            }
        }
        sink.flush();
        if (!sink.ok) failed[tid] = 1;
    };
    if (nt == 1 || n_reads < 64) {
        work(0, n_reads, 0);
    } else {
        std::vector<std::thread> th;
        // split by bytes, not by reads, so threads carry equal copy volume
        uint32_t lo = 0;
        for (int t = 0; t < nt; ++t) {
            uint64_t goal = off[n_reads] * (uint64_t)(t + 1) / nt;
            uint32_t hi = (uint32_t)(std::upper_bound(off.begin(), off.end(), goal) - off.begin());
            hi = std::min<uint32_t>(std::max<uint32_t>(hi, lo), n_reads);
            if (t == nt - 1) hi = n_reads;
            if (hi > lo) th.emplace_back(work, lo, hi, t);
            lo = hi;
        }
        for (auto& x : th) x.join();
    }
    for (char f : failed)
        if (f) return NS_EINVAL;
    return (int64_t)off[n_reads];
}
}  // namespace

int64_t ns_format_records(const uint8_t* seq, const uint8_t* qual, const NsReadMeta* reads, uint32_t n_reads,
                          const char* names, const uint64_t* name_off, int fastq, char* out, uint64_t out_cap,
                          int n_threads) {
    return format_records_impl(seq, qual, reads, n_reads, names, name_off, fastq, out, out_cap, n_threads, -1, 0);
}

int64_t ns_write_records(int fd, uint64_t file_off, const uint8_t* seq, const uint8_t* qual, const NsReadMeta* reads,
                         uint32_t n_reads, const char* names, const uint64_t* name_off, int fastq, int n_threads) {
    if (fd < 0) return NS_EINVAL;
    return format_records_impl(seq, qual, reads, n_reads, names, name_off, fastq, nullptr, 0, n_threads, fd, file_off);
}

// ---------------------------------------------------------------------------------------------------------
// host-side <out>_aligned_error_profile rows (mutate_read's log, simulator.py:2006-2008; header written by the caller):
// for every aligned segment, its error events right to left: name, position in the segment's reference, type, length,
// reference bases, read bases.  Events come from the segment's EVENT script (after the -k filter); when the
// homopolymer pass rewrote the emitted script, the bases of an event are the ones that pass fixed (hp_kernel.cuh:
// byte t of Philox-7 block (event index << 8) + (t >> 4) of stream ST_EMIT_B), otherwise they are read back from the
// sequence.  Two-call protocol like ns_format_records.
// ---------------------------------------------------------------------------------------------------------
namespace {
inline int dec_len(uint64_t v) {
    int n = 1;
    while (v >= 10) {
        v /= 10;
        ++n;
    }
    return n;
}
inline char* put_dec(char* p, uint64_t v) {
    char tmp[24];
    int n = 0;
    do {
        tmp[n++] = (char)('0' + v % 10);
        v /= 10;
    } while (v);
    while (n) *p++ = tmp[--n];
    return p;
}
struct EvRow {
    uint32_t type, len, ref_start, out_start, index, piece, ref_base;
    bool rewritten;
};
}  // namespace

static int64_t format_error_profile_impl(const uint8_t* seq, const NsReadMeta* reads, const NsPieceMeta* pieces, const uint32_t* ops,
                                         uint32_t n_reads, const uint8_t* ref_bases, const uint64_t* chrom_off, const char* names,
                                         const uint64_t* name_off, uint64_t seed, uint64_t first_id, char* out, uint64_t out_cap,
                                         int n_threads, int fd, uint64_t file_off) {
    if (!seq || !reads || !pieces || !ops || !ref_bases || !chrom_off || !names || !name_off) return NS_EINVAL;
    static const char kTypes[3][4] = {"mis", "ins", "del"};
    uint8_t comp[256];
    for (int c = 0; c < 256; ++c) comp[c] = (uint8_t)c;
    comp['A'] = 'T'; comp['T'] = 'A'; comp['C'] = 'G'; comp['G'] = 'C';
    const uint2 key = make_uint2((uint32_t)seed, (uint32_t)(seed >> 32));
    // one read: returns the bytes its rows take; writes them when p != nullptr
    auto do_read = [&](uint32_t i, char* p) -> uint64_t {
        const NsReadMeta& r = reads[i];
        const char* nm = names + name_off[i];
        const size_t nl = strlen(nm);
        const uint64_t rid = first_id + i;
        const uint32_t L = r.seq_len;
        const uint8_t* rs = seq + r.seq_off;
        const bool rev = r.reversed != 0;
        uint64_t bytes = 0;
        std::vector<EvRow> ev;
        // the pieces of one mutate_read call: a segment plus the pieces that continue it (NS_PIECE_CONT, intron retention)
        auto flush = [&]() {
            for (size_t e = ev.size(); e-- > 0;) {
                const EvRow& w = ev[e];
                const NsPieceMeta& pc = pieces[r.piece_first + w.piece];
                const uint64_t shown = (uint64_t)w.ref_base + w.ref_start;
                const uint64_t row = nl + 1 + dec_len(shown) + 1 + 3 + 1 + dec_len(w.len) + 1 + (uint64_t)w.len + 1 + w.len + 1;
                bytes += row;
                if (!p) continue;
                const uint64_t cstart = chrom_off[pc.chrom], clen = chrom_off[pc.chrom + 1] - cstart;
                const bool back = (pc.kind & NS_PIECE_REF_REV) != 0;
                memcpy(p, nm, nl);
                p += nl;
                *p++ = '\t';
                p = put_dec(p, shown);
                *p++ = '\t';
                memcpy(p, kTypes[w.type - 1], 3);
                p += 3;
                *p++ = '\t';
                p = put_dec(p, w.len);
                *p++ = '\t';
                char* refp = p;
                if (w.type == NS_OP_INS) {
                    memset(p, '-', w.len);
                } else {
                    for (uint32_t t = 0; t < w.len; ++t) {
                        const uint32_t f = w.ref_start + t;             // offset in the piece, in the direction of the read
                        uint64_t ab = (uint64_t)pc.pos + (back ? pc.ref_len - 1 - f : f);
                        if (ab >= clen) ab -= clen;
                        uint8_t c = ref_bases[cstart + ab];
                        if (c >= 'a' && c <= 'z') c = (uint8_t)(c - 32);
                        p[t] = (char)(back ? comp[c] : c);
                    }
                }
                p += w.len;
                *p++ = '\t';
                if (w.type == NS_OP_DEL) {
                    memset(p, '-', w.len);
                } else if (w.rewritten) {
                    for (uint32_t t = 0; t < w.len; ++t) {
                        const uint4 blk = philox4x32_7(make_uint4((uint32_t)rid, (uint32_t)(rid >> 32), stream_word(ST_EMIT_B, 0, w.piece),
                                                                  (w.index << 8) + (t >> 4)), key);
                        const uint32_t word = ((t >> 2) & 3u) == 0 ? blk.x : (((t >> 2) & 3u) == 1 ? blk.y : (((t >> 2) & 3u) == 2 ? blk.z : blk.w));
                        const uint32_t r8 = (word >> (8u * (t & 3u))) & 0xffu;
                        uint32_t bi;
                        if (w.type == NS_OP_INS) {
                            bi = r8 & 3u;
                        } else {
                            const char rc = refp[t];
                            const uint32_t orig = rc == 'C' ? 1u : (rc == 'T' ? 2u : (rc == 'G' ? 3u : 0u));
                            const uint32_t rr = r8 == 255u ? 0u : r8;
                            bi = (orig + 1u + rr % 3u) & 3u;
                        }
                        p[t] = "ACTG"[bi];
                    }
                } else {
                    for (uint32_t t = 0; t < w.len; ++t) {
                        const uint32_t x = w.out_start + t;
                        p[t] = (char)(rev ? comp[rs[L - 1 - x]] : rs[x]);
                    }
                }
                p += w.len;
                *p++ = '\n';
            }
            ev.clear();
        };
        uint32_t ref_base = 0;
        for (uint32_t k = 0; k < r.n_pieces; k += 2) {
            const NsPieceMeta& pc = pieces[r.piece_first + k];
            if (NS_PIECE_KIND(pc.kind) != NS_PIECE_SEGMENT) continue;
            if (!(pc.kind & NS_PIECE_CONT)) {
                flush();
                ref_base = 0;
            }
            const uint32_t* sc = ops + pc.ev_off;
            const bool rewritten = pc.ev_off != pc.op_off;
            uint32_t o = pc.out_rel, rf = 0;
            for (uint32_t j = 0; j < pc.ev_n_ops; ++j) {
                const uint32_t op = sc[j], ty = NS_OP_TYPE(op), ln = NS_OP_LEN(op);
                if (ty >= NS_OP_MIS && ty <= NS_OP_DEL && ln) ev.push_back(EvRow{ty, ln, rf, o, j, k, ref_base, rewritten});
                if (ty != NS_OP_DEL) o += ln;
                if (ty == NS_OP_COPY || ty == NS_OP_MIS || ty == NS_OP_DEL) rf += ln;
            }
            ref_base += pc.ref_len;
        }
        flush();
        return bytes;
    };
    int nt = std::max(1, std::min(n_threads, 64));
    std::vector<uint64_t> off((size_t)n_reads + 1, 0);
    {
        auto count = [&](uint32_t lo, uint32_t hi) {
            for (uint32_t i = lo; i < hi; ++i) off[i + 1] = do_read(i, nullptr);
        };
        if (nt == 1 || n_reads < 64) {
            count(0, n_reads);
        } else {
            std::vector<std::thread> th;
            for (int t = 0; t < nt; ++t) {
                uint32_t lo = (uint32_t)((uint64_t)n_reads * t / nt), hi = (uint32_t)((uint64_t)n_reads * (t + 1) / nt);
                if (hi > lo) th.emplace_back(count, lo, hi);
            }
            for (auto& x : th) x.join();
        }
        for (uint32_t i = 0; i < n_reads; ++i) off[i + 1] += off[i];
    }
    if (fd < 0) {
        if (!out) return (int64_t)off[n_reads];
        if (off[n_reads] > out_cap) return NS_ENOMEM;
    }
    std::vector<char> failed((size_t)nt + 1, 0);
    int next_tid = 0;
    auto fill = [&](uint32_t lo, uint32_t hi, int tid) {
        ChunkSink sink(out, fd, file_off);
        for (uint32_t i = lo; i < hi; ++i)
            if (off[i + 1] > off[i]) do_read(i, sink.reserve(off[i], (size_t)(off[i + 1] - off[i])));
        sink.flush();
        if (!sink.ok) failed[tid] = 1;
    };
    if (nt == 1 || n_reads < 64) {
        fill(0, n_reads, 0);
    } else {
        std::vector<std::thread> th;
        uint32_t lo = 0;
        for (int t = 0; t < nt; ++t) {
            uint64_t goal = off[n_reads] * (uint64_t)(t + 1) / nt;
            uint32_t hi = (uint32_t)(std::upper_bound(off.begin(), off.end(), goal) - off.begin());
            hi = std::min<uint32_t>(std::max<uint32_t>(hi, lo), n_reads);
            if (t == nt - 1) hi = n_reads;
            if (hi > lo) th.emplace_back(fill, lo, hi, next_tid++);
            lo = hi;
        }
        for (auto& x : th) x.join();
    }
    for (char f : failed)
        if (f) return NS_EINVAL;
    return (int64_t)off[n_reads];
}

int64_t ns_format_error_profile(const uint8_t* seq, const NsReadMeta* reads, const NsPieceMeta* pieces, const uint32_t* ops,
                                uint32_t n_reads, const uint8_t* ref_bases, const uint64_t* chrom_off, const char* names,
                                const uint64_t* name_off, uint64_t seed, uint64_t first_id, char* out, uint64_t out_cap,
                                int n_threads) {
    return format_error_profile_impl(seq, reads, pieces, ops, n_reads, ref_bases, chrom_off, names, name_off, seed, first_id, out,
                                     out_cap, n_threads, -1, 0);
}

int64_t ns_write_error_profile(int fd, uint64_t file_off, const uint8_t* seq, const NsReadMeta* reads, const NsPieceMeta* pieces,
                               const uint32_t* ops, uint32_t n_reads, const uint8_t* ref_bases, const uint64_t* chrom_off,
                               const char* names, const uint64_t* name_off, uint64_t seed, uint64_t first_id, int n_threads) {
    if (fd < 0) return NS_EINVAL;
    return format_error_profile_impl(seq, reads, pieces, ops, n_reads, ref_bases, chrom_off, names, name_off, seed, first_id, nullptr,
                                     0, n_threads, fd, file_off);
}


// ---------------------------------------------------------------------------------------------------------
// host-side read names (simulator.py:1390-1402 genome, :965-969 metagenome, :1188-1219 transcriptome, :1332-1343 perfect,
// :1511/:1529-1534 unaligned), written as NUL-terminated strings back to back -- the layout ns_format_records and
// ns_format_error_profile take.  flags: bit 0 perfect, bit 1 metagenome (gap lengths in the name), bit 2 transcriptome.
// ---------------------------------------------------------------------------------------------------------
int64_t ns_format_names(const NsReadMeta* reads, const NsPieceMeta* pieces, uint32_t n_reads, int kind, uint32_t flags,
                        uint64_t index_base, const char* chrom_names, const uint64_t* chrom_name_off, char* out,
                        uint64_t out_cap, uint64_t* name_off) {
    if (!reads || !pieces || !chrom_names || !chrom_name_off) return NS_EINVAL;
    const bool perfect = flags & 1u, meta = flags & 2u, trx = flags & 4u;
    // the reads are cut into ranges, one per thread; every thread builds the names of its range back to back in its own blob
    static const int name_threads = std::max(1, std::min(env_int("NANOSIM_B200_NAME_THREADS", 8), 64));
    const int nt = n_reads < 4096 ? 1 : name_threads;
    std::vector<std::string> blobs((size_t)nt);
    std::vector<std::vector<uint32_t>> lens((size_t)nt);
    auto work = [&](int tid) {
    const uint32_t lo = (uint32_t)((uint64_t)n_reads * tid / nt), hi = (uint32_t)((uint64_t)n_reads * (tid + 1) / nt);
    std::string& blob = blobs[tid];
    std::vector<uint32_t>& ln = lens[tid];
    blob.reserve((size_t)(hi - lo) * 64);
    ln.reserve(hi - lo);
    std::string nm;
    char num[32];
    auto add_num = [&](uint64_t v) {
        char* e = put_dec(num, v);
        nm.append(num, (size_t)(e - num));
    };
    for (uint32_t i = lo; i < hi; ++i) {
        const NsReadMeta& r = reads[i];
        const NsPieceMeta* pc = pieces + r.piece_first;
        const char strand = r.reversed ? 'R' : 'F';
        nm.clear();
        if (kind == NS_KIND_UNALIGNED) {
            nm += chrom_names + chrom_name_off[pc[0].chrom];
            nm += '_';
            add_num(pc[0].pos);
            nm += "_unaligned_";
            add_num(index_base + i);
            nm += '_';
            nm += strand;
            nm += "_0_";
            add_num(pc[0].ref_len);
            nm += "_0";
        } else if (trx && (pc[0].kind & NS_PIECE_GENOME)) {
            // intron-retention layout (:1188-1192, :1217-1219): transcript, genomic start of the first interval, the
            // retained-intron intervals the read covers in genomic order
            uint64_t first_pos = pc[0].pos, mid = 0;
            for (uint32_t k = 0; k < r.n_pieces; k += 2) {
                first_pos = std::min<uint64_t>(first_pos, pc[k].pos);
                mid += pc[k].ref_len;
            }
            nm += chrom_names + chrom_name_off[pc[0].ref_req];
            nm += '_';
            add_num(first_pos);
            nm += "_aligned_";
            add_num(index_base + i);
            bool any = false;
            for (uint32_t k = 0; k < r.n_pieces; k += 2) any = any || (pc[k].kind & NS_PIECE_RETAINED);
            if (any) {
                nm += "_RetainedIntron_";
                std::vector<std::pair<uint64_t, uint64_t>> ivs;              // in genomic order, whatever the strand
                for (uint32_t k = 0; k < r.n_pieces; k += 2)
                    if (pc[k].kind & NS_PIECE_RETAINED) ivs.emplace_back(pc[k].pos, (uint64_t)pc[k].pos + pc[k].ref_len);
                std::stable_sort(ivs.begin(), ivs.end());
                for (const auto& iv : ivs) {
                    add_num(iv.first);
                    nm += '-';
                    add_num(iv.second);
                    nm += ';';
                }
            }
            nm += '_';
            nm += strand;
            nm += '_';
            add_num(r.head);
            nm += '_';
            add_num(mid);
            nm += '_';
            add_num((uint64_t)r.tail + pc[0].polya_len);
        } else if (trx) {
            nm += chrom_names + chrom_name_off[pc[0].chrom];
            nm += '_';
            add_num(pc[0].pos);
            nm += perfect ? "_perfect_" : "_aligned_";
            add_num(index_base + i);
            nm += '_';
            nm += strand;
            nm += '_';
            add_num(r.head);
            nm += '_';
            add_num(pc[0].ref_len);
            nm += '_';
            add_num((uint64_t)r.tail + pc[0].polya_len);
        } else if (perfect) {
            uint64_t sum = 0;
            for (uint32_t k = 0; k < r.n_pieces; k += 2) {
                nm += chrom_names + chrom_name_off[pc[k].chrom];
                nm += '_';
                add_num(pc[k].pos);
                sum += pc[k].ref_len;
            }
            nm += "_perfect_";
            add_num(index_base + i);
            nm += '_';
            nm += strand;
            nm += "_0_";
            add_num(sum);
            nm += "_0";
        } else {
            for (uint32_t k = 0; k < r.n_pieces; ++k) {
                if (k & 1u) {
                    if (!meta) continue;
                    nm += ";gap_";
                    add_num(pc[k].out_len);
                    continue;
                }
                if (k) nm += ';';
                nm += chrom_names + chrom_name_off[pc[k].chrom];
                nm += '_';
                add_num(pc[k].pos);
            }
            nm += "_aligned_";
            add_num(index_base + i);
            if (r.n_pieces > 1) nm += "_chimeric";
            nm += '_';
            nm += strand;
            nm += '_';
            add_num(r.head);
            nm += '_';
            for (uint32_t k = 0; k < r.n_pieces; k += 2) {
                if (k) nm += ';';
                add_num(pc[k].ref_len);
            }
            nm += '_';
            add_num(r.tail);
        }
        blob.append(nm.c_str(), nm.size() + 1);
        ln.push_back((uint32_t)nm.size() + 1);
    }
    };
    if (nt == 1) {
        work(0);
    } else {
        std::vector<std::thread> th;
        for (int t = 1; t < nt; ++t) th.emplace_back(work, t);
        work(0);
        for (auto& x : th) x.join();
    }
    uint64_t total = 0;
    for (const std::string& bl : blobs) total += bl.size();
    if (!out) return (int64_t)total;
    if (total > out_cap) return NS_ENOMEM;
    uint64_t pos = 0;
    uint32_t i = 0;
    for (int t = 0; t < nt; ++t) {
        memcpy(out + pos, blobs[t].data(), blobs[t].size());
        if (name_off)
            for (uint32_t l : lens[t]) {
                name_off[i++] = pos;
                pos += l;
            }
        else
            pos += blobs[t].size();
    }
    return (int64_t)total;
}

// ---------------------------------------------------------------------------------------------------------
// FASTA / FASTQ reader of read_profile (simulator.py:341-349 with readfq :709-740): the file is mmap()ed and cut into
// line-aligned chunks; every thread finds the record headers of its chunk and counts its sequence bytes (pass 1), a prefix
// sum places the chunks, and the threads copy their sequence lines behind one another (pass 2).  Bytes are kept as they are
// (case, IUPAC codes); line ends (\n, \r\n) are dropped.  A FASTQ file (first byte '@') is read by one thread: its
// quality lines can begin with '>' or '@'.
// Two-call protocol: with bases == NULL only *n_records, *n_bases and *header_bytes are set.  rec_off gets n_records + 1
// offsets into bases; headers gets the header lines (without the marker) NUL-terminated back to back, header_off their starts.
// ---------------------------------------------------------------------------------------------------------
namespace {
struct FaChunk {
    const char* lo;
    const char* hi;
    uint64_t n_bases = 0;
    std::vector<std::pair<const char*, uint64_t>> heads;      // header line start (at the marker), sequence bytes of the chunk before it
};
inline const char* line_end(const char* p, const char* end) {
    const char* nl = (const char*)memchr(p, '\n', (size_t)(end - p));
    return nl ? nl : end;
}
inline size_t trimmed(const char* p, const char* e) {         // line length without a trailing \r
    return (e > p && e[-1] == '\r') ? (size_t)(e - p - 1) : (size_t)(e - p);
}
}  // namespace

int64_t ns_read_fasta(const char* path, uint8_t* bases, uint64_t bases_cap, uint64_t* rec_off, char* headers, uint64_t headers_cap,
                      uint64_t* header_off, uint32_t* n_records, uint64_t* n_bases, uint64_t* header_bytes, int n_threads) {
    if (!path || !n_records || !n_bases || !header_bytes) return NS_EINVAL;
    const int fd = open(path, O_RDONLY);
    if (fd < 0) return NS_EINVAL;
    struct stat sb;
    if (fstat(fd, &sb) != 0) {
        close(fd);
        return NS_EINVAL;
    }
    const size_t size = (size_t)sb.st_size;
    *n_records = 0;
    *n_bases = *header_bytes = 0;
    if (size == 0) {
        close(fd);
        return 0;
    }
    const char* base = (const char*)mmap(nullptr, size, PROT_READ, MAP_PRIVATE, fd, 0);
    close(fd);
    if (base == MAP_FAILED) return NS_ENOMEM;
    madvise((void*)base, size, MADV_SEQUENTIAL);
    const char* end = base + size;
    const bool fastq = base[0] == '@';
    int nt = fastq ? 1 : std::max(1, std::min(n_threads, 64));
    if (size < (size_t(1) << 22)) nt = 1;
    std::vector<FaChunk> ch((size_t)nt);
    for (int t = 0; t < nt; ++t) {                             // line-aligned chunk boundaries
        const char* p = base + size * (size_t)t / (size_t)nt;
        if (t > 0) {
            p = line_end(p - 1, end);
            if (p < end) ++p;
        }
        ch[t].lo = p;
        if (t > 0) ch[t - 1].hi = p;
    }
    ch[nt - 1].hi = end;
    const bool fill = bases != nullptr;
    // pass 1 / pass 2 over one chunk; FASTQ: sequence lines run to the '+' line, then as many quality bytes are skipped
    auto walk = [&](FaChunk& c, uint8_t* dst) {
        const char* p = c.lo;
        uint64_t count = 0;
        bool in_qual = false;
        uint64_t qual_left = 0, rec_bases = 0;
        while (p < c.hi) {
            const char* e = line_end(p, c.hi);
            const size_t len = trimmed(p, e);
            if (fastq && in_qual) {
                if (qual_left <= len) in_qual = false; else qual_left -= len;
            } else if (len && (p[0] == '>' || (fastq && p[0] == '@'))) {
                if (!dst) c.heads.emplace_back(p, count);
                rec_bases = 0;
            } else if (fastq && len && p[0] == '+') {
                in_qual = rec_bases > 0;
                qual_left = rec_bases;
            } else if (len) {
                if (dst) memcpy(dst + count, p, len);
                count += len;
                rec_bases += len;
            }
            p = e < c.hi ? e + 1 : c.hi;
        }
        if (!dst) c.n_bases = count;
    };
    {
        std::vector<std::thread> th;
        for (int t = 1; t < nt; ++t) th.emplace_back([&, t] { walk(ch[t], nullptr); });
        walk(ch[0], nullptr);
        for (auto& x : th) x.join();
    }
    uint64_t total = 0, n_rec = 0, hbytes = 0;
    std::vector<uint64_t> chunk_off((size_t)nt);
    for (int t = 0; t < nt; ++t) {
        chunk_off[t] = total;
        total += ch[t].n_bases;
        n_rec += ch[t].heads.size();
        for (auto& h : ch[t].heads) hbytes += trimmed(h.first, line_end(h.first, end));       // marker dropped, NUL added
    }
    *n_records = (uint32_t)n_rec;
    *n_bases = total;
    *header_bytes = hbytes;
    int64_t rc = (int64_t)total;
    if (fill) {
        if (total > bases_cap || hbytes > headers_cap || !rec_off || !headers || !header_off) {
            rc = NS_ENOMEM;
        } else {
            uint64_t r = 0, hpos = 0;
            for (int t = 0; t < nt; ++t)
                for (auto& h : ch[t].heads) {
                    rec_off[r] = chunk_off[t] + h.second;
                    const size_t hl = trimmed(h.first, line_end(h.first, end)) - 1;
                    header_off[r] = hpos;
                    memcpy(headers + hpos, h.first + 1, hl);
                    headers[hpos + hl] = 0;
                    hpos += hl + 1;
                    ++r;
                }
            rec_off[n_rec] = total;
            std::vector<std::thread> th;
            for (int t = 1; t < nt; ++t) th.emplace_back([&, t] { walk(ch[t], bases + chunk_off[t]); });
            walk(ch[0], bases + chunk_off[0]);
            for (auto& x : th) x.join();
        }
    }
    munmap((void*)base, size);
    return rc;
}

// ---------------------------------------------------------------------------------------------------------
// One NCCL broadcast of the reference at init (the reference's workers inherit the parent's seq_dict through fork(),
// simulator.py:1588-1622; ranks on different GPUs get it over NVLink instead of each parsing the FASTA).  NCCL is
// loaded at run time (dlopen), so the library has no link-time dependency on it.
// ---------------------------------------------------------------------------------------------------------
#if NS_HAVE_NCCL
namespace {
struct NcclApi {
    void* h = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*Broadcast)(const void*, void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    const char* (*GetErrorString)(ncclResult_t) = nullptr;
    bool ok = false;
};
NcclApi& nccl_api() {
    static NcclApi api = [] {
        NcclApi a;
        const char* override_path = getenv("NANOSIM_B200_NCCL_LIB");
        const char* names[] = {override_path, "libnccl.so.2", "libnccl.so"};
        for (const char* n : names) {
            if (!n || !*n) continue;
            a.h = dlopen(n, RTLD_NOW | RTLD_LOCAL);
            if (a.h) break;
        }
        if (!a.h) return a;
        a.GetUniqueId = (decltype(a.GetUniqueId))dlsym(a.h, "ncclGetUniqueId");
        a.CommInitRank = (decltype(a.CommInitRank))dlsym(a.h, "ncclCommInitRank");
        a.Broadcast = (decltype(a.Broadcast))dlsym(a.h, "ncclBroadcast");
        a.CommDestroy = (decltype(a.CommDestroy))dlsym(a.h, "ncclCommDestroy");
        a.GetErrorString = (decltype(a.GetErrorString))dlsym(a.h, "ncclGetErrorString");
        a.ok = a.GetUniqueId && a.CommInitRank && a.Broadcast && a.CommDestroy;
        return a;
    }();
    return api;
}
}  // namespace
#endif

int ns_get_reference(NsContext* ctx, uint8_t* bases, uint64_t cap) {
    if (!ctx || !bases) return NS_EINVAL;
    if (!ctx->have_ref) return fail(ctx, NS_ESTATE, "ns_get_reference: no reference set");
    if (cap < ctx->dref.genome_len) return fail(ctx, NS_ENOMEM, "ns_get_reference: buffer too small");
    CK(cudaSetDevice(ctx->device));
    CK(cudaMemcpy(bases, ctx->dref.bases, ctx->dref.genome_len, cudaMemcpyDeviceToHost));
    return NS_OK;
}

int ns_nccl_unique_id(uint8_t* id) {
#if NS_HAVE_NCCL
    if (!id) return NS_EINVAL;
    NcclApi& api = nccl_api();
    if (!api.ok) return NS_ESTATE;
    ncclUniqueId u;
    if (api.GetUniqueId(&u) != ncclSuccess) return NS_ECUDA;
    static_assert(sizeof(ncclUniqueId) == NS_NCCL_ID_BYTES, "ncclUniqueId size");
    memcpy(id, &u, sizeof u);
    return NS_OK;
#else
    (void)id;
    return NS_ESTATE;
#endif
}

int ns_bcast_nccl(NsContext* ctx, const uint8_t* id, int rank, int world, int root) {
    if (!ctx || !id || world < 1 || rank < 0 || rank >= world || root < 0 || root >= world) return fail(ctx, NS_EINVAL, "ns_bcast_nccl: bad argument");
#if NS_HAVE_NCCL
    if (ctx->borrowed) return fail(ctx, NS_ESTATE, "ns_bcast_nccl: a cloned context shares its parent's reference");
    if (rank == root && !ctx->have_ref) return fail(ctx, NS_ESTATE, "ns_bcast_nccl: the root rank sets its reference first (ns_set_reference)");
    if (world == 1) return NS_OK;
    NcclApi& api = nccl_api();
    if (!api.ok) return fail(ctx, NS_ESTATE, "ns_bcast_nccl: libnccl.so.2 could not be loaded (NANOSIM_B200_NCCL_LIB overrides the path)");
    CK(cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->stream;
    ncclUniqueId u;
    memcpy(&u, id, sizeof u);
    ncclComm_t comm = nullptr;
    ncclResult_t nr = api.CommInitRank(&comm, world, u, rank);
    if (nr != ncclSuccess) return fail(ctx, NS_ECUDA, "ns_bcast_nccl: ncclCommInitRank failed: %s", api.GetErrorString ? api.GetErrorString(nr) : "?");
    auto bc = [&](void* p, size_t bytes) { return bytes ? api.Broadcast(p, p, bytes, ncclUint8, root, comm, st) : ncclSuccess; };
    int rc = NS_OK;
    uint64_t* d_head = nullptr;
    void *t_bases = nullptr, *t_off = nullptr, *t_sp = nullptr, *t_circ = nullptr;
    do {
        uint64_t head[4] = {ctx->dref.genome_len, ctx->dref.n_chrom, ctx->dref.n_species, 0};
        if (cudaMalloc((void**)&d_head, sizeof head) != cudaSuccess) { rc = NS_ENOMEM; break; }
        cudaMemcpyAsync(d_head, head, sizeof head, cudaMemcpyHostToDevice, st);
        if (bc(d_head, sizeof head) != ncclSuccess) { rc = NS_ECUDA; break; }
        cudaMemcpyAsync(head, d_head, sizeof head, cudaMemcpyDeviceToHost, st);
        if (cudaStreamSynchronize(st) != cudaSuccess) { rc = NS_ECUDA; break; }
        const uint64_t n_bases = head[0];
        const uint32_t n_chrom = (uint32_t)head[1], n_species = (uint32_t)head[2];
        if (rank == root) {
            t_bases = ctx->ref_bases.p; t_off = ctx->ref_off.p; t_sp = ctx->ref_species.p; t_circ = ctx->ref_circular.p;
        } else {
            if (cudaMalloc(&t_bases, n_bases ? n_bases : 16) != cudaSuccess || cudaMalloc(&t_off, (n_chrom + 1) * sizeof(uint64_t)) != cudaSuccess ||
                (n_species && (cudaMalloc(&t_sp, n_chrom * 4) != cudaSuccess || cudaMalloc(&t_circ, n_chrom) != cudaSuccess))) { rc = NS_ENOMEM; break; }
        }
        if (bc(t_bases, n_bases) != ncclSuccess || bc(t_off, (n_chrom + 1) * sizeof(uint64_t)) != ncclSuccess ||
            (n_species && (bc(t_sp, (size_t)n_chrom * 4) != ncclSuccess || bc(t_circ, n_chrom) != ncclSuccess))) { rc = NS_ECUDA; break; }
        if (cudaStreamSynchronize(st) != cudaSuccess) { rc = NS_ECUDA; break; }
        if (rank != root) {
            NsReference r;
            r.bases = (const uint8_t*)t_bases;            // device pointers: ns_set_reference copies from them
            r.n_bases = n_bases;
            r.chrom_off = (const uint64_t*)t_off;
            r.n_chrom = n_chrom;
            r.n_species = n_species;
            r.chrom_species = (const uint32_t*)t_sp;
            r.chrom_circular = (const uint8_t*)t_circ;
            rc = ns_set_reference(ctx, &r);
        }
    } while (0);
    if (rank != root) {
        cudaFree(t_bases); cudaFree(t_off); cudaFree(t_sp); cudaFree(t_circ);
    }
    cudaFree(d_head);
    api.CommDestroy(comm);
    if (rc != NS_OK && ctx->err.empty()) ctx->err = "ns_bcast_nccl: broadcast failed";
    return rc;
#else
    (void)rank; (void)world; (void)root;
    return fail(ctx, NS_ESTATE, "ns_bcast_nccl: built without nccl.h");
#endif
}

}  // extern "C"
