// Launch-shape knobs.  The defaults are what the library ships with; the
// setters exist so one GPU session can sweep them (tools/tune.py) -- results
// are bit-reproducible only for a FIXED configuration, as documented for DOT.
#pragma once

namespace accblas {

struct Tuning {
    int dot_unroll = 4;        // 128-bit vectors of each operand in flight per thread
    int dot_ctas_per_sm = 0;   // grid = SMs * this (256 threads per CTA); 0 = all that are resident
    int dot_pdl = 1;           // programmatic dependent launch for back-to-back DOTs
    int gemv_unroll = 2;       // vectors per row in flight per lane
    int gemv_variant = 0;      // 0 = auto, 2 = CTA-per-2-rows, 3 = CTA-per-row, 4 = CTA-per-4-rows, 5 = CTA-per-8-rows
    int gemv_ctas_per_sm = 0;  // 0 = all row groups as separate CTAs
    int gemv_pipe = -1;        // default GEMV shape: -1 = per pair, 0 = plain loop, 1 = register pipeline, 3 / 4 = cp.async ring depth
    int gemv_intwords = 2;     // Acc<fp64,fp16>: words per 128-bit vector widened on the integer pipes (rest: F2F)
    int gemv_pdl = 1;          // programmatic dependent launch: back-to-back GEMVs overlap tail and ramp
    int gemv_force_pieces = 0; // 16-byte aligned operands through the 64-bit-load pipeline: 0 = Acc<fp64,fp16> only, 8 = all pairs, -1 = none
    int gemv_taper = 1;        // shorter row groups at the end of the grid
    int gemv_stages = 0;       // 0 = register path, 2..4 = bulk-copy ring depth
    int trsv_whole_block_spin = 1;  // TRSV: a caught-up CTA waits for a whole x block (1) or 32 entries at a time (0)
    int trsv_l2_ahead = 1024;  // TRSV: bytes per row of the groups of off-diagonal tiles requested into L2 one group ahead (0 = off)
};

Tuning& tuning();

}  // namespace accblas
