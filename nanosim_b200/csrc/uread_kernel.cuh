// Unaligned reads (simulation_unaligned, /root/reference/src/simulator.py:1482-1549), one WARP per read.
//
// unaligned_error_list (:1784-1830) draws an i.i.d. step per loop iteration (type with fixed cdf 0.4/0.7/0.85/1,
// length from the mixed models), so -- unlike the Markov chain of aligned reads -- the walk can be evaluated 32
// draws at a time: lane j takes draw (base + j), a warp prefix sum over the reference advance gives every draw its
// position, the first non-insertion draw that reaches m_ref ends the read exactly where the sequential loop would,
// and consecutive insertions are folded into the following step with a ballot/shuffle.  Draw k uses Philox block
// k+1 of the attempt's stream, i.e. exactly the words the sequential state machine in plan_kernel.cuh consumes for
// its k-th loop iteration, so both paths produce the same read lengths, rejections, strands and positions
// (tests/test_gpu_parity.py::test_unaligned_fast_path_equals_scripted_path).
//
// The WRITE pass lets each lane write the bases (and "unmapped"-state qualities, :1521) its own draw produces:
// with mutate_read's right-to-left string edits (:1957-1995) a step of length s at `pos` preceded by `a` inserted
// bases (key ceil(pos + 0.1) = pos + 1) becomes
//     match : ref[pos]                 + a random
//     mis   : sub(ref[pos])            + a random + sub(ref[pos+1 .. +rest]) + ref[.. covered]
//     del   :                            (a - covered) random               + ref[.. covered]
// with covered = min(a, s - 1) and rest = s - 1 - covered (the substitution / deletion also hits the inserted
// bases that sit inside its span; re-randomised random bases stay uniform).
#pragma once
#include "device_common.cuh"
#include "emit_kernel.cuh"
#include "plan_kernel.cuh"

struct UreadArgs {
    DevModel m;
    DevRef ref;
    DevCfg cfg;
    uint64_t first_id;
    uint32_t n_reads;
    NsReadMeta* reads;
    NsPieceMeta* pieces;
    uint8_t* seq;            // WRITE only
    uint8_t* qual;
    const uint32_t* qlut;    // [5][QLUT_SIZE]
    const uint32_t* qcdf;
    uint32_t* counter;
};

#define UREAD_WARPS 8

template <bool WRITE, bool FASTQ>
__global__ void __launch_bounds__(UREAD_WARPS * 32) uread_kernel(const __grid_constant__ UreadArgs a) {
    __shared__ uint32_t lut[(WRITE && FASTQ) ? QLUT_SIZE : 1];
    __shared__ uint32_t dsc_all[WRITE ? UREAD_WARPS : 1][5][32];
    uint32_t (*dsc)[32] = dsc_all[WRITE ? (threadIdx.x >> 5) : 0];
    const DevModel& m = a.m;
    const DevCfg& cfg = a.cfg;
    const int lane = threadIdx.x & 31;
    const uint32_t lane_lt = (1u << lane) - 1u;
    if (WRITE && FASTQ) {
        for (int i = threadIdx.x; i < QLUT_SIZE; i += blockDim.x) lut[i] = a.qlut[4 * QLUT_SIZE + i];   // "unmapped"
        __syncthreads();
    }
    const uint2 key = make_uint2((uint32_t)cfg.seed, (uint32_t)(cfg.seed >> 32));

    for (;;) {
        uint32_t slot = 0;
        if (lane == 0) slot = atomicAdd(a.counter, 1u);
        slot = __shfl_sync(0xffffffffu, slot, 0);
        if (slot >= a.n_reads) break;
        const uint64_t rid = a.first_id + slot;
        const uint32_t id_lo = (uint32_t)rid, id_hi = (uint32_t)(rid >> 32);
        uint32_t attempt = WRITE ? a.reads[slot].attempts : 0u;
        NsReadMeta rm;
        NsPieceMeta pm;
        if (WRITE) {
            rm = a.reads[slot];
            pm = a.pieces[slot];
        }
        for (;;) {   // rejection loop (:1503, :1517); every lane runs it redundantly on warp-uniform values
            const uint32_t sw = stream_word(ST_ATT, NS_KIND_UNALIGNED, attempt);
            Rng r0;
            r0.init(cfg.seed, rid, sw);
            // block 0 of the attempt's stream; -med/-sd: np.random.lognormal(log(median), sd) (:1494-1495)
            const double x = cfg.median_len > 0.0 ? lognormal_draw(log(cfg.median_len), cfg.sd_len, r0)
                                                  : kde_draw(m.unaligned, r0);
            const int64_t mr = (int64_t)x;
            if (mr <= 0) {                                       // middle_ref < min_l (:1503)
                ++attempt;
                continue;
            }
            const uint32_t m_ref = (uint32_t)mr;
            uint32_t base = 0, pos_base = 0, carry_a = 0, out_base = 0, middle_ref = m_ref, n_draws = 0;
            int64_t l_new = (int64_t)m_ref;
            bool done = false;
            while (!done) {
                const uint4 r = philox4x32_10(make_uint4(id_lo, id_hi, sw, base + lane + 1u), key);
                const uint32_t kind = r.x < 1717986918u ? 0u : (r.x < 3006477107u ? 1u : (r.x < 3650722201u ? 2u : 3u));
                uint32_t s = 1;
                if (kind != 0) s = alias_draw(m, kind == 1 ? 1u : (kind == 2 ? 2u : 3u), r.y);
                const bool nonins = kind != 2;
                const uint32_t adv = nonins ? s : 0u;
                const uint32_t P = pos_base + warp_incl_scan(adv, lane);
                const uint32_t I = warp_incl_scan(nonins ? 0u : s, lane);
                const uint32_t stop_mask = __ballot_sync(0xffffffffu, nonins && P >= m_ref);
                const int jstop = stop_mask ? __ffs(stop_mask) - 1 : 32;
                const bool valid = lane <= jstop;
                const uint32_t nonins_mask = __ballot_sync(0xffffffffu, nonins);
                const uint32_t below = nonins_mask & lane_lt;
                const int pn = below ? 31 - __clz(below) : -1;
                const uint32_t I_pn = __shfl_sync(0xffffffffu, I, pn < 0 ? 0 : pn);
                const uint32_t a_ins = nonins ? (pn >= 0 ? I - I_pn : I + carry_a) : 0u;
                // bases this draw produces / signed length change
                uint32_t covered = 0, rest = 0, outn = 0;
                if (nonins && valid) {
                    if (kind == 0) {
                        outn = 1 + a_ins;
                    } else {
                        covered = a_ins < s - 1 ? a_ins : s - 1;
                        rest = s - 1 - covered;
                        outn = kind == 1 ? s + a_ins : a_ins;
                    }
                }
                const uint32_t O = out_base + warp_incl_scan(outn, lane) - outn;
                int32_t delta = 0;
                if (valid) delta = kind == 2 ? (int32_t)s : (kind == 3 ? -(int32_t)s : 0);
#pragma unroll
                for (int d = 16; d > 0; d >>= 1) delta += __shfl_xor_sync(0xffffffffu, delta, d);
                l_new += delta;

                if (WRITE) {
                    // ---- emission: publish the draws' descriptors, then the warp writes the block's output bases
                    //      cooperatively (output index -> owning draw by binary search over the output offsets)
                    const uint32_t tot = __shfl_sync(0xffffffffu, O + outn, 31) - out_base;
                    dsc[0][lane] = O - out_base;
                    dsc[1][lane] = P - s;                                   // reference offset of the step
                    dsc[2][lane] = kind;
                    dsc[3][lane] = kind == 3 ? a_ins - covered : a_ins;     // surviving inserted bases
                    dsc[4][lane] = rest;
                    __syncwarp();
                    const uint64_t cstart = a.ref.chrom_off[pm.chrom];
                    const uint64_t clen = a.ref.chrom_off[pm.chrom + 1] - cstart;
                    const uint8_t* __restrict__ cb = a.ref.bases + cstart;
                    const bool rev = rm.reversed != 0;
                    uint8_t* sq = a.seq + rm.seq_off;
                    uint8_t* qq = FASTQ ? a.qual + rm.seq_off : nullptr;
                    for (uint32_t i = lane; i < tot; i += 32) {
                        uint32_t l = 0, h = 32;                              // last draw whose output offset <= i
                        while (h - l > 1) {
                            uint32_t mid = (l + h) >> 1;
                            if (dsc[0][mid] <= i) l = mid; else h = mid;
                        }
                        const uint32_t t = i - dsc[0][l], R = dsc[1][l], kd = dsc[2][l], n_ins = dsc[3][l], rst = dsc[4][l];
                        const uint32_t n_first = kd == 3 ? 0u : 1u;          // ref[pos] itself (copied or substituted)
                        const uint32_t n_mid = kd == 1 ? rst : 0u;           // further substituted reference bases
                        const uint32_t o = out_base + i;
                        // one 32-bit word per base: Philox-7 block o>>2, word o&3 of the read's unaligned base stream
                        const uint4 w4 = philox4x32_7(make_uint4(id_lo, id_hi, stream_word(ST_EMIT_B, NS_KIND_UNALIGNED, 0), o >> 2), key);
                        const uint32_t w = (o & 2u) ? ((o & 1u) ? w4.w : w4.z) : ((o & 1u) ? w4.y : w4.x);
                        const uint32_t wn = (o & 2u) ? ((o & 1u) ? w4.x : w4.w) : ((o & 1u) ? w4.z : w4.y);   // next word of the block
                        const uint32_t r8 = w & 0xffu;                       // low byte: base choice; high 24 bits: quality
                        const uint32_t t3 = __umulhi(__byte_perm(w, wn, 0x0444), 3u);
                        uint32_t oi = r8 & 3u;
                        int roff = -1;
                        bool sub = false;
                        if (t < n_first) {
                            roff = 0;
                            sub = kd == 1;
                        } else if (t < n_first + n_ins) {
                            roff = -1;
                        } else if (t < n_first + n_ins + n_mid) {
                            roff = 1 + (int)(t - n_first - n_ins);
                            sub = true;
                        } else {
                            roff = 1 + (int)rst + (int)(t - n_first - n_ins - n_mid);
                        }
                        if (roff >= 0) {
                            uint64_t ab = (uint64_t)pm.pos + R + (uint32_t)roff;
                            if (ab >= clen) ab -= clen;
                            uint32_t c = __ldg(&cb[ab]);
                            if (c - 'a' < 26u) c -= 32;
                            if (!acgt_fast(c)) c = converted_ref_base(c, cfg.seed, rid, 0u, R + (uint32_t)roff);
                            oi = base_idx(c);
                            if (sub) oi = (oi + 1u + t3) & 3u;
                        }
                        const uint32_t dst = rev ? rm.seq_len - 1 - o : o;
                        sq[dst] = (uint8_t)emit_char(rev ? oi ^ 2u : oi, cfg.uracil);
                        if (FASTQ) {
                            const uint32_t e = lut[w >> (32 - QLUT_BITS)];
                            uint32_t q = qual_char_fast(e, w);
                            if (q & 0x80u) q = qual_char_exact(a.qcdf + 4 * NS_QUAL_SLOTS, w);
                            qq[dst] = (uint8_t)q;
                        }
                    }
                    __syncwarp();
                }

                if (jstop < 32) {
                    const uint32_t Pstop = __shfl_sync(0xffffffffu, P, jstop);
                    if (Pstop > middle_ref) {                      // overrun extends the segment (:1826-1828)
                        l_new += Pstop - middle_ref;
                        middle_ref = Pstop;
                    }
                    n_draws = base + (uint32_t)jstop + 1u;
                    done = true;
                } else {
                    pos_base = __shfl_sync(0xffffffffu, P, 31);
                    const uint32_t I31 = __shfl_sync(0xffffffffu, I, 31);
                    if (nonins_mask) {
                        const int last = 31 - __clz(nonins_mask);
                        carry_a = I31 - __shfl_sync(0xffffffffu, I, last);
                    } else {
                        carry_a += I31;
                    }
                    out_base = __shfl_sync(0xffffffffu, O + outn, 31);     // O already includes the old out_base
                    base += 32;
                }
            }
            if (WRITE) break;
            const bool ok = middle_ref >= cfg.min_len && middle_ref <= cfg.max_len && l_new >= (int64_t)cfg.min_len &&
                            l_new <= (int64_t)cfg.max_len;
            if (!ok) {
                ++attempt;
                continue;
            }
            // accepted: strand from the next block of the attempt stream (:1526-1527), position (extract_read)
            const uint4 rs = philox4x32_10(make_uint4(id_lo, id_hi, sw, n_draws + 1u), key);
            const uint32_t reversed = u01_double(((uint64_t)rs.x << 32) | rs.y) > (double)m.strandness;
            Rng pr;
            pr.init(cfg.seed, rid, stream_word(ST_POS, NS_KIND_UNALIGNED, attempt));
            uint32_t chrom = 0, ppos = 0;
            if (cfg.metagenome) draw_position_meta(a.ref, pr, -1, middle_ref, chrom, ppos);
            else if (cfg.transcriptome) draw_position_trx(a.ref, pr, middle_ref, chrom, ppos);
            else draw_position(a.ref, cfg, pr, middle_ref, chrom, ppos);
            if (lane == 0) {
                NsPieceMeta p;
                p.op_off = 0;
                p.n_ops = 0;
                p.kind = NS_PIECE_UNALIGNED;
                p.chrom = chrom;
                p.pos = ppos;
                p.ref_len = middle_ref;
                p.out_len = (uint32_t)l_new;
                p.out_rel = 0;
                p.l_new = (uint32_t)l_new;
                p.ref_req = m_ref;
                p.read_slot = slot;
                p.ev_off = 0;
                p.ev_n_ops = 0;
                p.polya_len = 0;
                a.pieces[slot] = p;
                NsReadMeta q;
                q.seq_off = 0;
                q.seq_len = (uint32_t)l_new;
                q.head = 0;
                q.tail = 0;
                q.piece_first = slot;
                q.n_pieces = 1;
                q.reversed = (uint8_t)reversed;
                q.flags = 0;
                q.attempts = attempt;
                a.reads[slot] = q;
            }
            break;
        }
    }
}
