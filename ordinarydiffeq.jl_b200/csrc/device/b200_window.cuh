// b200_window.cuh — a software-pipelining window for the flagged square roots of a user RHS that is inlined as one
// large basic block (device/b200_vern7_wide.cuh).
//
// ptxas schedules such a block greedily: it starts every independent root/quotient chain of the text at once (21 for
// Pleiades), their intermediates fill the register file, and the second half of the block is emitted as strictly
// serial 8-cycle dependent chains.  With one warp per scheduler nothing hides those.  The window bounds how far ahead a
// root may start: the argument of the i-th sqrt of the text is gated on the range flag as it stood B200_WIDE_WINDOW
// roots earlier (a select that only changes the value when the flag is already raised — and a raised flag discards the
// whole evaluation in favour of the plain operators), which makes the i-th chain data-dependent on the quotients that
// preceded the (i - W)-th root.  At most W + 1 groups are in flight; results are unchanged.
#pragma once
#include "b200_base.cuh"

#ifndef B200_WIDE_WINDOW
#define B200_WIDE_WINDOW 0
#endif

B200_D real b200_sqrt_window(real x, bool& bad, bool* win) {
#if B200_WIDE_WINDOW > 0
    x = win[B200_WIDE_WINDOW - 1] ? (real)1 : x;
#pragma unroll
    for (int k = B200_WIDE_WINDOW - 1; k > 0; --k) win[k] = win[k - 1];
    win[0] = bad;
#else
    (void)win;
#endif
    return b200_sqrt_fast(x, bad);
}
