// Emit kernel: applies each piece's edit script to the reference and writes bases + base qualities.
//
// This is extract_read's slice (/root/reference/src/simulator.py:1750-1781), case_convert (:743-755), mutate_read's
// string edits and quality interleaving (:1957-2015), the head/tail synthesis (:1421-1427), reverse_complement
// (:1433-1435, :1675-1680) and the per-state quality draws (model_base_qualities.py:120-130) in ONE pass:
//
//   * a warp owns one piece (an aligned segment with its head/tail, a chimeric gap, or an unaligned read) and
//     streams its ops 32 at a time; a warp-wide scan turns op lengths into (output start, reference start) pairs
//     kept in a shared-memory ring;
//   * every lane produces one 16-byte output chunk per step: it binary-searches the ring for the op covering its
//     first base, then walks ops/reference bytes sequentially, so global stores are 16-byte vectors, 512 B per
//     warp and fully coalesced for both the base and the quality stream;
//   * reverse-strand reads are produced directly in output order by walking the edit script and the reference
//     backwards and complementing (no second pass over the read);
//   * all randomness is Philox keyed by (seed, read id, chunk index), so the bytes do not depend on the batch,
//     the launch geometry or the GPU count.
#pragma once
#include "device_common.cuh"

#define EMIT_WARPS 8
#define EMIT_RING 256
#define QLUT_BITS 11
#define QLUT_SIZE (1 << QLUT_BITS)
#define QLUT_FRAC_BITS (24 - QLUT_BITS)     // quality uniforms are 24-bit

struct EmitArgs {
    DevRef ref;
    DevCfg cfg;
    uint32_t kind;
    uint64_t first_id;
    const NsReadMeta* reads;
    const NsPieceMeta* pieces;
    const uint32_t* ops;
    uint32_t n_pieces;
    uint8_t* seq;
    uint8_t* qual;
    const uint32_t* qlut;        // [5][QLUT_SIZE] packed bucket table (built on the host from qual_cdf)
    const uint32_t* qcdf;        // [5][94], 24-bit fixed point
    uint32_t* counter;
};

__device__ __forceinline__ uint32_t warp_incl_scan(uint32_t v, int lane) {
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        uint32_t t = __shfl_up_sync(0xffffffffu, v, d);
        if (lane >= d) v += t;
    }
    return v;
}

template <bool FASTQ>
__global__ void __launch_bounds__(EMIT_WARPS * 32) emit_kernel(const __grid_constant__ EmitArgs a) {
    extern __shared__ uint32_t smem[];
    uint32_t* lut = smem;                                        // FASTQ only
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint32_t* ring = smem + (FASTQ ? NS_N_QUAL_STATES * QLUT_SIZE : 0) + warp * (3 * EMIT_RING);
    uint32_t* ring_out = ring;
    uint32_t* ring_ref = ring + EMIT_RING;
    uint32_t* ring_op = ring + 2 * EMIT_RING;
    if (FASTQ) {
        for (int i = threadIdx.x; i < NS_N_QUAL_STATES * QLUT_SIZE; i += blockDim.x) lut[i] = a.qlut[i];
        __syncthreads();
    }

    for (;;) {
        uint32_t piece = 0;
        if (lane == 0) piece = atomicAdd(a.counter, 1u);
        piece = __shfl_sync(0xffffffffu, piece, 0);
        if (piece >= a.n_pieces) break;
        const NsPieceMeta pm = a.pieces[piece];
        if (pm.out_len == 0) continue;
        const NsReadMeta rm = a.reads[pm.read_slot];
        const uint64_t rid = a.first_id + pm.read_slot;
        const bool rev = rm.reversed != 0;
        const uint32_t out_len = pm.out_len, n_ops = pm.n_ops, ref_len = pm.ref_len;
        const uint32_t A = rev ? rm.seq_len - pm.out_rel - out_len : pm.out_rel;   // piece start in read coordinates
        const uint64_t cstart = a.ref.chrom_off[pm.chrom];
        const uint64_t clen = a.ref.chrom_off[pm.chrom + 1] - cstart;
        const uint8_t* __restrict__ cbase = a.ref.bases + cstart;
        const uint32_t* __restrict__ ops = a.ops + pm.op_off;
        uint8_t* seq_out = a.seq + rm.seq_off;
        uint8_t* qual_out = FASTQ ? a.qual + rm.seq_off : nullptr;
        const bool unmapped = pm.kind != NS_PIECE_SEGMENT;

        uint32_t t_loaded = 0, t_ret = 0, out_loaded = 0, ref_loaded = 0, prog = 0;
        while (prog < out_len) {
            // ---- 1. stream ops into the ring until the next 32 chunks are covered or the ring is full
            const uint32_t first_chunk = (A + prog) >> 4;
            uint32_t target = (first_chunk + 32) * 16 - A;
            if (target > out_len) target = out_len;
            while (out_loaded < target && t_loaded < n_ops && (t_loaded - t_ret) + 32 <= EMIT_RING) {
                uint32_t t = t_loaded + lane;
                uint32_t op = 0;
                if (t < n_ops) op = __ldg(&ops[rev ? n_ops - 1 - t : t]);
                uint32_t ty = op >> 28, len = op_len(op);
                uint32_t o = (ty == NS_OP_DEL) ? 0u : len;
                uint32_t r = (ty < 2u || ty == NS_OP_DEL) ? len : 0u;
                uint32_t so = warp_incl_scan(o, lane), sr = warp_incl_scan(r, lane);
                if (t < n_ops) {
                    ring_out[t % EMIT_RING] = out_loaded + so - o;
                    ring_ref[t % EMIT_RING] = ref_loaded + sr - r;
                    ring_op[t % EMIT_RING] = op;
                }
                out_loaded += __shfl_sync(0xffffffffu, so, 31);
                ref_loaded += __shfl_sync(0xffffffffu, sr, 31);
                t_loaded += (n_ops - t_loaded < 32u) ? n_ops - t_loaded : 32u;
            }
            __syncwarp();
            // ---- 2. what can be produced now: whole chunks up to the loaded frontier (or the piece end)
            uint32_t lim = out_loaded < target ? out_loaded : target;
            if (lim < out_len) lim = ((A + lim) & ~15u) > A + prog ? ((A + lim) & ~15u) - A : prog;
            // (ring holds >= 64 ops beyond t_ret, i.e. >= 1 whole chunk, so lim > prog whenever ops remain)
            // ---- 3. one 16-byte chunk per lane
            const uint32_t chunk = first_chunk + lane;
            uint32_t cs = chunk * 16;                                  // chunk start, read coordinates
            uint32_t lo = cs > A + prog ? cs : A + prog;
            uint32_t hi = cs + 16 < A + lim ? cs + 16 : A + lim;
            if (lo < hi) {
                const uint32_t plo = lo - A;                           // piece coordinates
                // upper_bound over ring_out in [t_ret, t_loaded): last op whose output start <= plo
                uint32_t l = t_ret, h = t_loaded;
                while (h - l > 1) {
                    uint32_t mid = (l + h) >> 1;
                    if (ring_out[mid % EMIT_RING] <= plo) l = mid; else h = mid;
                }
                uint32_t k = l;
                uint32_t op = ring_op[k % EMIT_RING];
                uint32_t ty = op >> 28;
                uint32_t within = plo - ring_out[k % EMIT_RING];
                uint32_t rem = op_len(op) - within;
                uint32_t rpos = ring_ref[k % EMIT_RING] + ((ty < 2u) ? within : 0u);

                // ---- all randomness of the chunk up front, position-indexed, identical in every lane's control flow:
                //      base i uses byte i of `bw` (substitution / inserted base / IUPAC member) and bits [24i, 24i+24)
                //      of `qw` (its quality value).
                const uint2 key = make_uint2((uint32_t)a.cfg.seed, (uint32_t)(a.cfg.seed >> 32));
                const uint32_t id_lo = (uint32_t)rid, id_hi = (uint32_t)(rid >> 32);
                const uint4 bw4 = philox4x32_7(make_uint4(id_lo, id_hi, stream_word(ST_EMIT_B, a.kind, chunk), 0u), key);
                const uint32_t bw[4] = {bw4.x, bw4.y, bw4.z, bw4.w};
                uint32_t qw[13];
                if (FASTQ) {
                    const uint32_t sw = stream_word(ST_EMIT_Q, a.kind, chunk);
#pragma unroll
                    for (int j = 0; j < 3; ++j) {
                        uint4 t = philox4x32_7(make_uint4(id_lo, id_hi, sw, (uint32_t)j), key);
                        qw[4 * j] = t.x; qw[4 * j + 1] = t.y; qw[4 * j + 2] = t.z; qw[4 * j + 3] = t.w;
                    }
                    qw[12] = 0;
                }
                const uint32_t flip = rev ? 2u : 0u;

                uint32_t sb[4] = {0, 0, 0, 0}, sq[4] = {0, 0, 0, 0};
                const uint32_t i0 = lo - cs, i1 = hi - cs;
#pragma unroll
                for (int i = 0; i < 16; ++i) {
                    if ((uint32_t)i >= i0 && (uint32_t)i < i1) {
                        while (rem == 0) {
                            ++k;
                            op = ring_op[k % EMIT_RING];
                            ty = op >> 28;
                            rem = (ty == NS_OP_DEL) ? 0u : op_len(op);
                            rpos = ring_ref[k % EMIT_RING];
                        }
                        --rem;
                        const uint32_t r8 = (bw[i >> 2] >> (8 * (i & 3))) & 0xffu;
                        const uint32_t r8n = (bw[((i + 1) & 15) >> 2] >> (8 * ((i + 1) & 3))) & 0xffu;
                        const uint32_t rr = (r8 == 255u) ? r8n : r8;           // 0..254 -> exactly uniform mod 3
                        const uint32_t t3 = rr - 3u * ((rr * 171u) >> 9);
                        uint32_t oi = r8 & 3u;                                 // random.choice(BASES) / np.random.choice
                        if (ty < 2u) {                                         // COPY or MIS: reads the reference
                            uint32_t f = rev ? ref_len - 1 - rpos : rpos;
                            uint64_t ab = (uint64_t)pm.pos + f;
                            if (ab >= clen) ab -= clen;                        // circular wrap (:1756-1760)
                            uint32_t c = __ldg(&cbase[ab]);
                            ++rpos;
                            if (c - 'a' < 26u) c -= 32;
                            if (!acgt_fast(c)) c = converted_ref_base(c, a.cfg.seed, rid, piece - rm.piece_first, f);
                            oi = base_idx(c);
                            if (ty == NS_OP_MIS) oi = (oi + 1u + t3) & 3u;     // one of the three other bases
                        } else if (ty == NS_OP_LIT) {
                            oi = (op >> 26) & 3u;                              // literal base of a rewritten homopolymer
                        }
                        sb[i >> 2] |= emit_char(oi ^ flip, a.cfg.uracil) << (8 * (i & 3));
                        if (FASTQ) {
                            // quality state: COPY->match(2) MIS->mis(0) INS->ins(1) HT->ht(3); gap/unaligned -> unmapped(4)
                            const uint32_t qs = unmapped ? 4u : (ty == NS_OP_LIT ? ((op >> 24) & 3u) : ((0x30102u >> (4u * ty)) & 7u));
                            const int bit = 24 * i;
                            const uint32_t u24 = __funnelshift_r(qw[bit >> 5], qw[(bit >> 5) + 1], bit & 31) & 0xffffffu;
                            const uint32_t e = lut[qs * QLUT_SIZE + (u24 >> QLUT_FRAC_BITS)];
                            uint32_t q = e & 0xffu;
                            if (e >> 31) {                                     // bucket spans >2 quality values: exact scan
                                const uint32_t* cdf = a.qcdf + qs * NS_QUAL_SLOTS;
                                while (q < NS_QUAL_SLOTS - 1 && u24 >= __ldg(&cdf[q])) ++q;
                            } else {
                                q += ((u24 & ((1u << QLUT_FRAC_BITS) - 1u)) >= ((e >> 8) & 0x3fffu)) ? 1u : 0u;
                            }
                            sq[i >> 2] |= (q + 33u) << (8 * (i & 3));
                        }
                    }
                }
                if (i0 == 0 && i1 == 16) {
                    *reinterpret_cast<uint4*>(seq_out + cs) = make_uint4(sb[0], sb[1], sb[2], sb[3]);
                    if (FASTQ) *reinterpret_cast<uint4*>(qual_out + cs) = make_uint4(sq[0], sq[1], sq[2], sq[3]);
                } else {
#pragma unroll
                    for (int i = 0; i < 16; ++i) {
                        if ((uint32_t)i >= i0 && (uint32_t)i < i1) {
                            seq_out[cs + i] = (uint8_t)(sb[i >> 2] >> (8 * (i & 3)));
                            if (FASTQ) qual_out[cs + i] = (uint8_t)(sq[i >> 2] >> (8 * (i & 3)));
                        }
                    }
                }
            }
            // ---- 4. retire ops that end before the new frontier
            prog = lim;
            {
                uint32_t l = t_ret, h = t_loaded;
                while (h - l > 1) {
                    uint32_t mid = (l + h) >> 1;
                    if (ring_out[mid % EMIT_RING] <= prog) l = mid; else h = mid;
                }
                t_ret = l;
            }
            __syncwarp();
        }
    }
}
