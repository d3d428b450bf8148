// Emit kernel: applies each piece's edit script to the reference and writes bases + base qualities.
//
// This is extract_read's slice (/root/reference/src/simulator.py:1750-1781), case_convert (:743-755), mutate_read's
// string edits and quality interleaving (:1957-2015), the head/tail synthesis (:1421-1427), reverse_complement
// (:1433-1435, :1675-1680) and the per-state quality draws (model_base_qualities.py:120-130) in ONE pass:
//
//   * a warp owns one piece (an aligned segment with its head/tail, a chimeric gap, or an unaligned read) and
//     streams its ops 32 at a time; a warp-wide scan turns them into ring entries {output start, absolute reference
//     offset, length, per-op constants} in shared memory (deletions only advance the reference offset and get no
//     entry; a pad entry aligns the piece to its first 16-byte chunk, a sentinel closes it);
//   * every lane produces one 16-byte output chunk per step: it binary-searches the ring for the entry covering its
//     first base, then runs a BRANCH-FREE, fully unrolled 16-base loop: a predicated ring advance, a predicated
//     reference byte load, one table lookup for case/IUPAC classification, one for the quality bucket.  Lanes of a
//     warp sit in different ops, so anything branchy would be executed by everybody anyway; the rare cases (IUPAC
//     codes, quality buckets that hold more than two values) only raise a flag and are patched afterwards;
//   * reverse-strand reads are produced directly in output order by walking the edit script and the reference
//     backwards with a complemented character table (no second pass over the read);
//   * all randomness is Philox keyed by (seed, read id, chunk index): one 32-bit word per base (8 bits base choice,
//     24 bits quality uniform), so the bytes do not depend on the batch, the launch geometry or the GPU count.
#pragma once
#include "device_common.cuh"

#define EMIT_WARPS 8
#define EMIT_RING 256
#ifndef EMIT_MIN_BLOCKS
#define EMIT_MIN_BLOCKS 3     // resident blocks per SM the register allocation aims for
#endif

#define EMIT_F_REF 4u        // info bit: the base is read from the reference (COPY, MIS)
#define EMIT_F_MIS 8u        // info bit: ... and substituted by one of the three other bases

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
    const uint32_t* order;       // piece processing order (longest first) or null
};

__device__ __forceinline__ uint32_t warp_incl_scan(uint32_t v, int lane) {
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        uint32_t t = __shfl_up_sync(0xffffffffu, v, d);
        if (lane >= d) v += t;
    }
    return v;
}

// Per-op constants of a ring entry: [1:0] 3 when the base is random (INS, HT, pad), [2] EMIT_F_REF, [3] EMIT_F_MIS,
// [15:8] the character classified when the reference is not read ('A' -> index 0, or the literal base of a LIT op),
// [31:16] byte offset of the quality state's bucket table.
// ref_comp: the piece classifies reference bytes with the COMPLEMENTING table (NS_PIECE_REF_REV); the characters stored for
// ops that do not read the reference are then pre-complemented so that they classify to the intended base.
__device__ __forceinline__ uint32_t emit_op_info(uint32_t op, bool unmapped, bool ref_comp = false) {
    const uint32_t ty = op >> 28;
    // quality state: COPY->match(2) MIS->mis(0) INS->ins(1) HT->ht(3) LIT->its own; gap/unaligned -> unmapped(4)
    const uint32_t qs = unmapped ? 4u : (ty == NS_OP_LIT ? ((op >> 24) & 3u) : ((0x30102u >> (4u * ty)) & 7u));
    uint32_t info = (qs * QLUT_SIZE * 4u) << 16;
    const uint32_t zero = ref_comp ? (uint32_t)'T' : (uint32_t)'A';          // classifies to base index 0
    if (ty < 2u) info |= EMIT_F_REF | (ty == NS_OP_MIS ? EMIT_F_MIS : 0u) | (zero << 8);
    else if (ty == NS_OP_LIT) info |= idx_base(((op >> 26) & 3u) ^ (ref_comp ? 2u : 0u)) << 8;
    else info |= 3u | (zero << 8);
    return info;
}

struct EmitPiece {
    const uint8_t* cbase;    // first base of the chromosome
    uint32_t pos, clen;      // piece start within the chromosome, chromosome length
    uint32_t rdir;           // the reference is walked backwards (reverse read XOR minus-strand piece)
    uint32_t ref_comp;       // reference bases are complemented on classification (minus-strand piece)
    uint32_t tbl;            // output characters of base indices 0..3 (complemented for reverse reads)
    uint32_t piece_in_read;
    uint64_t rid;
};

// Exact slow path for ONE base (IUPAC reference codes, quality buckets with more than two values): locates the ring
// entry of padded piece coordinate x again and redoes the base with the same random word.  Returns char | qchar << 8.
template <bool FASTQ>
__device__ __noinline__ uint32_t emit_fix_base(const EmitArgs& a, const EmitPiece& pc, const uint4* ring, uint32_t w_ret,
                                               uint32_t w_loaded, uint32_t x, uint32_t w, uint32_t p) {
    uint32_t l = w_ret, h = w_loaded;
    while (h - l > 1) {
        uint32_t mid = (l + h) >> 1;
        if (ring[mid & (EMIT_RING - 1)].x <= x) l = mid; else h = mid;
    }
    const uint4 e = ring[l & (EMIT_RING - 1)];
    const uint32_t within = x - e.x, info = e.w;
    uint32_t oi;
    if (info & EMIT_F_REF) {
        uint32_t rabs = pc.rdir ? e.y - within : e.y + within;
        if (rabs >= pc.clen) rabs += pc.rdir ? pc.clen : 0u - pc.clen;
        uint32_t c = __ldg(pc.cbase + rabs);
        if (c - 'a' < 26u) c -= 32;
        if (!acgt_fast(c)) {
            const uint32_t f = rabs >= pc.pos ? rabs - pc.pos : rabs + pc.clen - pc.pos;
            c = converted_ref_base(c, a.cfg.seed, pc.rid, pc.piece_in_read, f);
        }
        oi = base_idx(c) ^ (pc.ref_comp ? 2u : 0u);
        if (info & EMIT_F_MIS) oi = (oi + 1u + __umulhi(p, 3u)) & 3u;
    } else {
        oi = (info & 3u) ? (w & 3u) : (base_idx((info >> 8) & 0xffu) ^ (pc.ref_comp ? 2u : 0u));
    }
    uint32_t out = (pc.tbl >> (8u * oi)) & 0xffu;
    if (FASTQ) {
        const uint32_t qs = (info >> 16) / (QLUT_SIZE * 4u);
        const uint32_t qe = __ldg(&a.qlut[qs * QLUT_SIZE + (w >> (32 - QLUT_BITS))]);
        out |= ((qe & 0x80u) ? qual_char_exact(a.qcdf + qs * NS_QUAL_SLOTS, w, qe) : qual_char_fast(qe, w)) << 8;
    }
    return out;
}

// The branch-free 16-base walk of one chunk (see the header comment).  WRAPS: the piece crosses the origin of a circular
// chromosome (:1756-1760).
template <bool FASTQ, bool WRAPS, bool COMP>
__device__ __forceinline__ void emit_chunk16(const uint4* ring, const uint32_t* lut, const uint8_t* cvt_tables,
                                             const uint8_t* __restrict__ cbase, const uint32_t (&W)[FASTQ ? 16 : 4], uint32_t& k,
                                             uint32_t& rem, uint32_t& rabs, uint32_t& info, uint32_t dir, uint32_t clen,
                                             uint32_t wrap_fix, uint32_t tbl, uint32_t (&sb)[4], uint32_t (&sq)[4], uint32_t& bad,
                                             uint32_t& slow) {
    const uint8_t* cvt = cvt_tables + (COMP ? 256 : 0);    // compile-time offset: the lookup stays [register + immediate]
    uint32_t sel = 0;
#pragma unroll
    for (int i = 0; i < 16; ++i) {
        if (rem == 0) {                                    // next entry (never empty, never a deletion)
            ++k;
            const uint4 e = ring[k & (EMIT_RING - 1)];
            rem = e.z;
            rabs = e.y;
            info = e.w;
        }
        --rem;
        // base i's random bits: rb low 2 bits = random.choice(BASES); rp = 16+ uniform bits, MSB-aligned, whose
        // product with 3 picks one of the three other bases (bias 2^-16)
        const uint32_t rb = FASTQ ? W[i] : W[i >> 2] >> (8 * (i & 3));
        const uint32_t rp = FASTQ ? __byte_perm(W[i], W[(i + 1) & 15], 0x0444)
                                  : __byte_perm(W[i >> 2], W[((i + 1) & 15) >> 2], ((i & 3) << 12) | ((4 + ((i + 1) & 3)) * 0x111));
        uint32_t c = (info >> 8) & 0xffu;
        if (info & EMIT_F_REF) {
            c = __ldg(cbase + rabs);
            rabs += dir;
            if (WRAPS && rabs >= clen) rabs += wrap_fix;
        }
        const uint32_t code = cvt[c];
        bad |= code;
        uint32_t v = code | (rb & info & 3u);
        if (info & EMIT_F_MIS) v += 1u + __umulhi(rp, 3u);
        sel += v << (4 * (i & 3));
        if (FASTQ) {
            const uint32_t w = W[i];
            const uint32_t e = *reinterpret_cast<const uint32_t*>(reinterpret_cast<const char*>(lut) + (info >> 16) +
                                                                  ((w >> 19) & 0x1ffcu));
            slow |= e;
            sq[i >> 2] |= qual_char_fast(e, w) << (8 * (i & 3));
        }
        if ((i & 3) == 3) {
            sb[i >> 2] = __byte_perm(tbl, 0, sel & 0x3333u);
            sel = 0;
        }
    }
}

template <bool FASTQ>
__global__ void __launch_bounds__(EMIT_WARPS * 32, EMIT_MIN_BLOCKS) emit_kernel(const __grid_constant__ EmitArgs a) {
    extern __shared__ uint4 smem4[];
    uint32_t* lut = reinterpret_cast<uint32_t*>(smem4);                          // FASTQ only
    uint8_t* cvt_tables = reinterpret_cast<uint8_t*>(lut + (FASTQ ? NS_N_QUAL_STATES * QLUT_SIZE : 0));
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint4* ring = reinterpret_cast<uint4*>(cvt_tables + 512) + warp * EMIT_RING;
    if (FASTQ)
        for (int i = threadIdx.x; i < NS_N_QUAL_STATES * QLUT_SIZE; i += blockDim.x) lut[i] = a.qlut[i];
    {   // case_convert classification: base index 0..3 (A C T G), 4 = needs the IUPAC path
        uint32_t c = threadIdx.x;
        uint32_t u = (c - 'a' < 26u) ? c - 32u : c;
        cvt_tables[c] = (uint8_t)(acgt_fast(u) ? base_idx(u) : 4u);
        cvt_tables[256 + c] = (uint8_t)(acgt_fast(u) ? base_idx(u) ^ 2u : 4u);     // complementing twin (NS_PIECE_REF_REV)
    }
    __syncthreads();
    const uint2 key = make_uint2((uint32_t)a.cfg.seed, (uint32_t)(a.cfg.seed >> 32));
    const uint32_t lane_lt = (1u << lane) - 1u;

    for (;;) {
        uint32_t piece = 0;
        if (lane == 0) piece = atomicAdd(a.counter, 1u);
        piece = __shfl_sync(0xffffffffu, piece, 0);
        if (piece >= a.n_pieces) break;
        if (a.order) piece = __ldg(&a.order[piece]);
        const NsPieceMeta pm = a.pieces[piece];
        if (pm.out_len == 0) continue;
        const NsReadMeta rm = a.reads[pm.read_slot];
        const bool rev = rm.reversed != 0;
        // a minus-strand piece (intron retention, ns_reemit) reads the genome backwards and complemented: together with the
        // read's own orientation that fixes the walking direction, and the complement moves into the classification table
        const bool ref_comp = (pm.kind & NS_PIECE_REF_REV) != 0;
        const bool rdir = rev != ref_comp;
        const uint32_t n_ops = pm.n_ops, ref_len = pm.ref_len;
        const uint32_t A = rev ? rm.seq_len - pm.out_rel - pm.out_len : pm.out_rel;    // piece start in read coordinates
        const uint32_t pad = A & 15u, P0 = A - pad;          // x = read coordinate - P0: chunks are x/16
        const uint32_t x_end = pad + pm.out_len;
        const uint64_t cstart = a.ref.chrom_off[pm.chrom];
        EmitPiece pc;
        pc.cbase = a.ref.bases + cstart;
        pc.pos = pm.pos;
        pc.clen = (uint32_t)(a.ref.chrom_off[pm.chrom + 1] - cstart);
        pc.rdir = rdir;
        pc.ref_comp = ref_comp;
        {
            const uint32_t t = a.cfg.uracil ? 0x47554341u : 0x47544341u;             // "ACTG" / "ACUG"
            pc.tbl = rev ? __byte_perm(t, 0, 0x1032) : t;                            // complement: A<->T, C<->G
        }
        pc.piece_in_read = piece - rm.piece_first;
        pc.rid = a.first_id + pm.read_slot;
        const uint8_t* __restrict__ cbase = pc.cbase;
        const uint32_t clen = pc.clen, tbl = pc.tbl;
        const bool wraps = (uint64_t)pm.pos + ref_len > clen;                        // circular wrap (:1756-1760)
        const uint32_t dir = rdir ? 0xffffffffu : 1u, wrap_fix = rdir ? clen : 0u - clen;
        const uint32_t* __restrict__ ops = a.ops + pm.op_off;
        uint8_t* seq_out = a.seq + rm.seq_off + P0;
        uint8_t* qual_out = FASTQ ? a.qual + rm.seq_off + P0 : nullptr;
        const bool unmapped = NS_PIECE_KIND(pm.kind) != NS_PIECE_SEGMENT;
        const uint32_t id_lo = (uint32_t)pc.rid, id_hi = (uint32_t)(pc.rid >> 32);
        const uint32_t info_pad = emit_op_info(NS_OP_HT << 28, unmapped, ref_comp);

        uint32_t t_loaded = 0, w_loaded = 0, w_ret = 0, out_loaded = pad, ref_loaded = 0, prog = 0;
        bool closed = false;
        if (pad) {
            if (lane == 0) ring[0] = make_uint4(0u, 0u, pad, info_pad);
            w_loaded = 1;
        }
        while (prog < x_end) {
            // ---- 1. stream ops into the ring until the next 32 chunks are covered or the ring is full
            uint32_t target = prog + 512u < x_end ? prog + 512u : x_end;
            while (out_loaded < target && t_loaded < n_ops && (w_loaded - w_ret) + 33u <= EMIT_RING) {
                const uint32_t t = t_loaded + lane;
                uint32_t op = 0;
                if (t < n_ops) op = __ldg(&ops[rev ? n_ops - 1 - t : t]);
                const uint32_t ty = op >> 28, len = op_len(op);
                const uint32_t o = (ty == NS_OP_DEL) ? 0u : len;
                const uint32_t r = (ty < 2u || ty == NS_OP_DEL) ? len : 0u;
                const uint32_t so = warp_incl_scan(o, lane), sr = warp_incl_scan(r, lane);
                const bool keep = o != 0u;
                const uint32_t bal = __ballot_sync(0xffffffffu, keep);
                if (keep) {
                    const uint32_t rstart = ref_loaded + sr - r;                      // reference bases consumed before this op
                    uint32_t ab = pm.pos + (rdir ? ref_len - 1u - rstart : rstart);   // first base in walking order
                    if (wraps && ab >= clen) ab -= clen;
                    ring[(w_loaded + __popc(bal & lane_lt)) & (EMIT_RING - 1)] =
                        make_uint4(out_loaded + so - o, ab, len, emit_op_info(op, unmapped, ref_comp));
                }
                out_loaded += __shfl_sync(0xffffffffu, so, 31);
                ref_loaded += __shfl_sync(0xffffffffu, sr, 31);
                w_loaded += __popc(bal);
                t_loaded += (n_ops - t_loaded < 32u) ? n_ops - t_loaded : 32u;
            }
            if (!closed && (t_loaded == n_ops || out_loaded >= x_end)) {
                // sentinel: walking past the end of the piece stays inside the ring (ops left over are deletions)
                if (lane == 0) ring[w_loaded & (EMIT_RING - 1)] = make_uint4(out_loaded, 0u, 0x7fffffffu, info_pad);
                ++w_loaded;
                closed = true;
            }
            __syncwarp();
            // ---- 2. what can be produced now: whole chunks up to the loaded frontier (or the piece end)
            uint32_t lim = out_loaded < target ? out_loaded : target;
            if (lim < x_end) lim &= ~15u;
            // (the ring holds >= 200 entries of >= 1 base beyond w_ret, so lim > prog whenever ops remain)
            // ---- 3. one 16-byte chunk per lane
            const uint32_t cs = prog + 16u * lane;
            uint32_t k = w_ret;
            if (cs < lim) {
                uint32_t l = w_ret, h = w_loaded;     // last entry whose output start <= cs
                while (h - l > 1) {
                    uint32_t mid = (l + h) >> 1;
                    if (ring[mid & (EMIT_RING - 1)].x <= cs) l = mid; else h = mid;
                }
                k = l;
                uint32_t rem, rabs, info;
                {
                    const uint4 e = ring[k & (EMIT_RING - 1)];
                    const uint32_t within = cs - e.x;
                    rem = e.z - within;
                    info = e.w;
                    rabs = rdir ? e.y - within : e.y + within;
                    if (wraps && rabs >= clen) rabs += wrap_fix;
                }
                // ---- all randomness of the chunk up front, position-indexed: base i owns one 32-bit word (FASTQ: low
                //      byte = base choice, high 24 bits = quality uniform) or one byte (FASTA)
                const uint32_t chunk = (P0 + cs) >> 4;
                uint32_t W[FASTQ ? 16 : 4];
                if (FASTQ) {
                    const uint32_t sw = stream_word(ST_EMIT_Q, a.kind, chunk);
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const uint4 t = philox4x32_7(make_uint4(id_lo, id_hi, sw, (uint32_t)j), key);
                        W[4 * j] = t.x; W[4 * j + 1] = t.y; W[4 * j + 2] = t.z; W[4 * j + 3] = t.w;
                    }
                } else {
                    const uint4 t = philox4x32_7(make_uint4(id_lo, id_hi, stream_word(ST_EMIT_B, a.kind, chunk), 0u), key);
                    W[0] = t.x; W[1] = t.y; W[2] = t.z; W[3] = t.w;
                }
                uint32_t sb[4], sq[4] = {0, 0, 0, 0};
                uint32_t bad = 0, slow = 0;
                // three instantiations: plain, circular wrap, complementing (minus-strand genome pieces; keeps the wrap check)
                if (ref_comp) emit_chunk16<FASTQ, true, true>(ring, lut, cvt_tables, cbase, W, k, rem, rabs, info, dir, clen, wrap_fix, tbl, sb, sq, bad, slow);
                else if (wraps) emit_chunk16<FASTQ, true, false>(ring, lut, cvt_tables, cbase, W, k, rem, rabs, info, dir, clen, wrap_fix, tbl, sb, sq, bad, slow);
                else emit_chunk16<FASTQ, false, false>(ring, lut, cvt_tables, cbase, W, k, rem, rabs, info, dir, clen, wrap_fix, tbl, sb, sq, bad, slow);
                // ---- rare exact paths, one base at a time
                if ((bad & 4u) || (FASTQ && (slow & 0x80u))) {
#pragma unroll
                    for (int i = 0; i < 16; ++i) {
                        const bool need = (bad & 4u) || (FASTQ && ((sq[i >> 2] >> (8 * (i & 3))) & 0x80u));
                        if (need) {
                            const uint32_t rb = FASTQ ? W[i] : W[i >> 2] >> (8 * (i & 3));
                            const uint32_t rp = FASTQ ? __byte_perm(W[i], W[(i + 1) & 15], 0x0444)
                                                      : __byte_perm(W[i >> 2], W[((i + 1) & 15) >> 2],
                                                                    ((i & 3) << 12) | ((4 + ((i + 1) & 3)) * 0x111));
                            const uint32_t f = emit_fix_base<FASTQ>(a, pc, ring, w_ret, w_loaded, cs + i, rb, rp);
                            const uint32_t m = 0xffu << (8 * (i & 3));
                            sb[i >> 2] = (sb[i >> 2] & ~m) | ((f & 0xffu) << (8 * (i & 3)));
                            if (FASTQ) sq[i >> 2] = (sq[i >> 2] & ~m) | ((f >> 8) << (8 * (i & 3)));
                        }
                    }
                }
                if (cs >= pad && cs + 16u <= x_end) {
                    *reinterpret_cast<uint4*>(seq_out + cs) = make_uint4(sb[0], sb[1], sb[2], sb[3]);
                    if (FASTQ) *reinterpret_cast<uint4*>(qual_out + cs) = make_uint4(sq[0], sq[1], sq[2], sq[3]);
                } else {
#pragma unroll
                    for (int i = 0; i < 16; ++i) {
                        if (cs + i >= pad && cs + i < x_end) {
                            seq_out[cs + i] = (uint8_t)(sb[i >> 2] >> (8 * (i & 3)));
                            if (FASTQ) qual_out[cs + i] = (uint8_t)(sq[i >> 2] >> (8 * (i & 3)));
                        }
                    }
                }
            }
            // ---- 4. retire: the entry the last active lane stopped in starts before the new frontier
            {
                const uint32_t n_act = (lim - prog + 15u) >> 4;
                const uint32_t k_last = __shfl_sync(0xffffffffu, k, n_act ? n_act - 1u : 0u);
                if (n_act) w_ret = k_last;
            }
            prog = lim;
            __syncwarp();
        }
    }
}
