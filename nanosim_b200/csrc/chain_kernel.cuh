// The error chain of LONG aligned segments (error_list, /root/reference/src/simulator.py:1833-1916), one WARP per segment.
//
// plan_kernel walks a read's Markov chain with one lane: ~9 bases per event, a few hundred cycles of dependent latency per
// event (a table lookup keyed by the previous match length, then the next).  For the bulk of a batch that is hidden by
// the other lanes; the longest reads (100-600 kb) are not: the kernel ends when the lane that drew the longest read has
// finished its ~20 000 - 60 000 events, milliseconds after everything else.  This pre-pass takes those segments away
// from it.
//
// What makes the chain serial is small: the state going into an event is (error state, bin of the previous match length),
// 7 x n_bins values, while the event's random words are known up front (Philox block k + 1 of the segment's stream).  So the
// warp first evaluates, 32 events at a time, one per lane, EVERY outcome the event could have: the error type for each of
// the 7 error states (comparisons), the error length for each of the 3 types and the next match length for each of the
// n_bins bins (3 + n_bins independent alias-table lookups, all in flight together), each with the bin it leads to.  The
// candidates go to shared memory; the serial part that is left is a walk over the 32 events in which one event costs one
// shared-memory read that depends on the previous one.  Positions, the stopping rule and the op offsets are warp scans.
//
// The ops, the counts and the lengths are exactly those of plan_kernel's own walk (same stream, same blocks), which still
// handles later attempts and replays of these segments; tests compare the two (NS_FLAG_EMIT_WHOLE switches this pre-pass off).
#pragma once
#include "device_common.cuh"
#include "emit_kernel.cuh"
#include "plan_kernel.cuh"

#define CHAIN_MIN_LEN 32768u          // drawn reference length above which a segment is walked here
#define CHAIN_WARPS 4
#define CHAIN_MISS 0xffffffffu

struct ChainShared {                  // per warp; a value is (length << 5) | bin(length)
    uint32_t mt[NS_MAX_BINS][32];     // next match length if the previous one fell into bin b (CHAIN_MISS: the ECDF miss slot)
    uint32_t st[3][32];               // error length if the error is mis / ins / del
    uint32_t ec[32];                  // error type for each of the 7 error states, 2 bits each (0: keep the previous type)
};

struct AChain {
    uint32_t n_ops, middle_ref, out_len, l_new;
};

// ops go to ops[0 .. cap) (counted beyond that)
__device__ __forceinline__ AChain aligned_chain_warp(const DevModel& m, const uint8_t* bin_lut, uint2 key, uint32_t id_lo, uint32_t id_hi,
                                                     uint32_t sw, uint32_t m_ref, uint32_t* ops, uint32_t cap, ChainShared& cs, int lane) {
    AChain res;
    uint32_t n_ops = 0, out_len = 0;
    // first match from _first_match.hist (:1843-1850); no extension when it overshoots
    const uint32_t fm = alias_draw(m, 0, philox4x32_10(make_uint4(id_lo, id_hi, sw, 0u), key).x);
    {
        const uint32_t first = fm < m_ref ? fm : m_ref;
        if (first) {
            if (lane == 0 && cap > 0) ops[0] = (NS_OP_COPY << 28) | first;
            n_ops = 1;
            out_len = first;
        }
    }
    if (fm >= m_ref) {
        res.n_ops = n_ops;
        res.middle_ref = m_ref;
        res.out_len = out_len;
        res.l_new = m_ref;
        return res;
    }
    const uint32_t nb = m.n_bins;
    const uint32_t one = (1u << 5) | match_bin(m, bin_lut, 1u);
    uint32_t pos_base = fm, prev_bin = match_bin(m, bin_lut, fm), err_state = 0, last_err = 3;
    bool pz = fm == 0;                        // previous match length was 0
    int64_t dl = 0;                           // inserted - deleted bases
    bool carry_iz = false;                    // the previous block ended with an insertion followed by no match
    uint32_t carry_ins_len = 0;
    for (uint32_t base = 0;; base += 32u) {
        // ---- every outcome of event base + lane
        {
            const uint4 r = philox4x32_10(make_uint4(id_lo, id_hi, sw, base + (uint32_t)lane + 1u), key);
            uint32_t ecv = 0;
#pragma unroll
            for (int s = 0; s < NS_N_ERR_STATES; ++s) {
                const uint32_t e = r.x < m.trans[s][0] ? 1u : (r.x < m.trans[s][1] ? 2u : (r.x >= m.trans[s][2] ? 3u : 0u));
                ecv |= e << (2 * s);
            }
            cs.ec[lane] = ecv;
#pragma unroll
            for (uint32_t e = 1; e <= 3; ++e) {
                const uint32_t len = alias_draw(m, e, r.y);
                cs.st[e - 1][lane] = (len << 5) | match_bin(m, bin_lut, len);
            }
            for (uint32_t b = 0; b < nb; ++b) {
                const uint32_t v = alias_draw(m, 4 + b, r.z);
                cs.mt[b][lane] = (v == m.tab_n[4 + b] - 1u) ? CHAIN_MISS : ((v << 5) | match_bin(m, bin_lut, v));
            }
        }
        __syncwarp();
        // ---- the serial walk: which outcome it was
        uint32_t my_e = 1, my_step = 0, my_mt = 0;
        for (int j = 0; j < 32; ++j) {
            uint32_t c = cs.mt[prev_bin][j];
            uint32_t e = (cs.ec[j] >> (2u * err_state)) & 3u;
            if (!e) e = last_err;                            // dead gap of the (1-p_del, 1) interval: stale value
            last_err = e;
            const uint32_t sv = cs.st[e - 1u][j];
            if (c == CHAIN_MISS) c = sv;                      // ECDF miss: `step` keeps the error length (:1895-1898)
            if (pz && (c >> 5) == 0u) c = one;
            if (lane == j) {
                my_e = e;
                my_step = sv >> 5;
                my_mt = c >> 5;
            }
            pz = (c >> 5) == 0u;
            err_state = e + (pz ? 3u : 0u);                   // prev_error += "0" (:1913-1914)
            prev_bin = c & 31u;
        }
        __syncwarp();
        // ---- positions and the stopping rule: the first event after which pos >= m_ref
        const uint32_t P = pos_base + warp_incl_scan((my_e != 2u ? my_step : 0u) + my_mt, lane);
        const uint32_t stop_mask = __ballot_sync(0xffffffffu, P >= m_ref);
        const int jstop = stop_mask ? __ffs(stop_mask) - 1 : 32;
        const bool valid = lane <= jstop;
        // ---- ops: the error, then the match (dropped when empty).  An insertion that follows an insertion with no match in
        //      between replaces it (e_dict[pos - 0.5] is overwritten, :1882): it takes its predecessor's place in the script
        const bool is_ins = my_e == 2u;
        const bool ins_zero = is_ins && my_mt == 0u;
        const uint32_t iz_mask = __ballot_sync(0xffffffffu, ins_zero), ins_mask = __ballot_sync(0xffffffffu, is_ins);
        const bool prev_iz = lane ? ((iz_mask >> (lane - 1)) & 1u) != 0u : carry_iz;
        const bool replacing = is_ins && prev_iz;
        const bool replaced = ins_zero && lane < 31 && lane < jstop && ((ins_mask >> (lane + 1)) & 1u);
        const uint32_t cnt = valid ? ((replacing ? 0u : 1u) + (my_mt ? 1u : 0u)) : 0u;
        const uint32_t incl = warp_incl_scan(cnt, lane);
        const uint32_t at = n_ops + incl - cnt;
        if (valid) {
            const uint32_t se = replacing ? at - 1u : at;
            const uint32_t ty = my_e == 1u ? NS_OP_MIS : (is_ins ? NS_OP_INS : NS_OP_DEL);
            if (!replaced && se < cap) ops[se] = (ty << 28) | my_step;
            if (my_mt && se + 1u < cap) ops[se + 1u] = (NS_OP_COPY << 28) | my_mt;
        }
        // ---- lengths
        int32_t d = 0;
        uint32_t o = 0;
        if (valid) {
            d = is_ins ? (int32_t)my_step : (my_e == 3u ? -(int32_t)my_step : 0);
            o = my_mt + (my_e == 1u ? my_step : 0u) + ((is_ins && !replaced) ? my_step : 0u);
        }
#pragma unroll
        for (int k = 16; k > 0; k >>= 1) {
            d += __shfl_xor_sync(0xffffffffu, d, k);
            o += __shfl_xor_sync(0xffffffffu, o, k);
        }
        dl += d;
        out_len += o;
        if ((ins_mask & 1u) && carry_iz) out_len -= carry_ins_len;      // lane 0 replaced the insertion the last block ended with
        n_ops += __shfl_sync(0xffffffffu, incl, 31);
        if (jstop < 32) {
            const uint32_t Pstop = __shfl_sync(0xffffffffu, P, jstop);
            res.n_ops = n_ops;
            res.middle_ref = Pstop;                            // >= m_ref: the overrun extends the segment (:1826-1828 analogue, :1876-1880, :1904-1908)
            res.out_len = out_len;
            res.l_new = (uint32_t)((int64_t)Pstop + dl);
            return res;
        }
        pos_base = __shfl_sync(0xffffffffu, P, 31);
        carry_iz = ((iz_mask >> 31) & 1u) != 0u;
        carry_ins_len = __shfl_sync(0xffffffffu, my_step, 31);
    }
}

struct ChainArgs {
    DevModel m;
    DevCfg cfg;
    uint32_t kind;
    uint64_t first_id;
    uint32_t n_reads;
    const uint32_t* n_seg;          // per read (nullptr => 1)
    const uint32_t* piece_first;    // per read (nullptr => read index)
    NsPieceMeta* pieces;
    uint32_t* ops;
    const uint32_t* order;          // read slots by decreasing total drawn length
    uint32_t* counter;
    const uint32_t* abort;
    uint32_t min_len;               // segments whose drawn reference length exceeds this are walked here
};

// Results travel in the piece record: n_ops, ref_len (middle_ref), out_len, l_new, polya_len = 1 as "walked" mark (the field is
// only used in transcriptome mode, which this pre-pass does not serve).  The ops of a read's FIRST piece start one word into
// the slot: the head op, whose length the plan kernel draws, goes in front.
__global__ void __launch_bounds__(CHAIN_WARPS * 32) chain_kernel(const __grid_constant__ ChainArgs a) {
    if (a.abort && *a.abort) return;
    __shared__ uint8_t bin_lut[BIN_LUT_SIZE];
    __shared__ ChainShared cs_all[CHAIN_WARPS];
    for (uint32_t i = threadIdx.x; i < BIN_LUT_SIZE; i += blockDim.x) bin_lut[i] = (uint8_t)match_bin_scan(a.m, i);
    __syncthreads();
    const int lane = threadIdx.x & 31;
    ChainShared& cs = cs_all[threadIdx.x >> 5];
    const uint2 key = make_uint2((uint32_t)a.cfg.seed, (uint32_t)(a.cfg.seed >> 32));
    for (;;) {
        uint32_t idx = 0;
        if (lane == 0) idx = atomicAdd(a.counter, 1u);
        idx = __shfl_sync(0xffffffffu, idx, 0);
        if (idx >= a.n_reads) break;
        const uint32_t slot = a.order ? a.order[idx] : idx;
        const uint32_t ns = a.n_seg ? a.n_seg[slot] : 1u;
        const uint32_t pf = a.piece_first ? a.piece_first[slot] : slot;
        const uint64_t rid = a.first_id + slot;
        uint64_t total = 0;
        for (uint32_t q = 0; q < 2u * ns - 1u; ++q) total += a.pieces[pf + q].ref_req;
        if (total <= a.min_len) break;                 // reads come by decreasing total length: nothing long is left
        for (uint32_t q = 0; q < 2u * ns - 1u; q += 2u) {
            NsPieceMeta& pm = a.pieces[pf + q];
            const uint32_t m_ref = pm.ref_req;
            if (m_ref <= a.min_len) continue;
            const uint32_t shift = q == 0 ? 1u : 0u;
            const uint64_t op_off = pm.op_off;
            const uint32_t room = (uint32_t)(a.pieces[pf + q + 1].op_off - op_off);
            const AChain ch = aligned_chain_warp(a.m, bin_lut, key, (uint32_t)rid, (uint32_t)(rid >> 32),
                                                 stream_word(ST_CHAIN, a.kind, q & 31u), m_ref, a.ops + op_off + shift,
                                                 room > shift ? room - shift : 0u, cs, lane);
            if (lane == 0) {
                pm.n_ops = ch.n_ops;
                pm.ref_len = ch.middle_ref;
                pm.out_len = ch.out_len;
                pm.l_new = ch.l_new;
                pm.polya_len = 1;
            }
        }
    }
}
