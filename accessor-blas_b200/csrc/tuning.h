// Launch-shape knobs.  The defaults are what the library ships with; the
// setters exist so one GPU session can sweep them (tools/tune.py) -- results
// are bit-reproducible only for a FIXED configuration, as documented for DOT.
// accblas_tune() validates every value against the range given in capi.cu.
#pragma once

namespace accblas {

struct Tuning {
    int dot_unroll = 0;        // 128-bit vectors of each operand in flight per thread (1 / 2 / 4); 0 = default (4)
    int dot_block = 0;         // threads per CTA (256 / 512 / 1024); 0 = default (1024 for Acc<fp64,fp64>, else 256)
    int dot_ctas_per_sm = 0;   // grid = SMs * this; 0 = all that are resident
    int dot_pdl = 1;           // programmatic dependent launch for back-to-back DOTs
    int dot_pool_pct = 12;     // DOT: share of the tiles (%) handed out dynamically in chunks at the end (0 = static partition only)
    int dot_chunk_tiles = 4;   // DOT: tiles per dynamically assigned chunk (grows with n so that there are at most 4096 chunks)
    int dot_intmix = 0;        // Acc<fp64,fp32>: widen x on the integer pipes (experiment)
    int gemv_unroll = 2;       // vectors per row in flight per lane
    int gemv_variant = 0;      // 0 = auto, 2 = CTA-per-2-rows, 3 = CTA-per-row, 4 = CTA-per-4-rows, 5 = CTA-per-8-rows
    int gemv_ctas_per_sm = 0;  // 0 = all row groups as separate CTAs
    int gemv_pipe = -1;        // default GEMV shape: -1 = per pair, 0 = plain loop, 1 = register pipeline, 3 / 4 = cp.async ring depth
    int gemv_intwords = 2;     // Acc<fp64,fp16>: words per 128-bit vector widened on the integer pipes (rest: F2F)
    int gemv_pdl = 1;          // programmatic dependent launch: back-to-back GEMVs overlap tail and ramp
    int gemv_force_pieces = 0; // 16-byte aligned operands through the 64-bit-load pipeline: 0 = Acc<fp64,fp16> only, 8 = all pairs, -1 = none
    int gemv_taper = 1;        // shorter row groups at the end of the grid
    int gemv_stages = 0;       // 0 = register path, 2..4 = bulk-copy ring depth
    int trsv_variant = -1;     // -1 = per storage type (measured on B200: the cluster kernel wins for fp16 storage, 12-24 %; the single-CTA kernel for fp32 / fp64 storage), 0 = cluster kernel (DSMEM hand-off), 1 = one CTA per block row through L2
    int trsv_whole_block_spin = 1;  // TRSV variant 1: a caught-up CTA waits for a whole x block (1) or 32 entries at a time (0)
    int trsv_l2_ahead = -1;    // TRSV: bytes per row of the groups of off-diagonal tiles requested into L2 one group ahead (0 = off; -1 = per pair, see trsv_default_l2_ahead)
    int fill_generic = 0;      // fill_uniform: 1 = per-row kernel with __ddiv_rn also for contiguous outputs (A/B check)
};

Tuning& tuning();

}  // namespace accblas
