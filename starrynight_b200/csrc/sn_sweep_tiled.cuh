// sn_sweep_tiled.cuh -- TMA-staged shared-memory tile kernel (fast path).  Placeholder until the
// kernel lands: reports "unsupported" so every lattice runs on the colour-pass kernel.
#pragma once
#include "sn_common.cuh"

bool sn_tiled_supported(const sn_handle *h, std::string *why)
{
    (void)h;
    if (why) *why = "tiled kernel not built";
    return false;
}
int sn_tiled_prepare(sn_handle *) { return SN_OK; }
void sn_tiled_release(sn_handle *) {}
int sn_sweep_tiled_launch(sn_handle *, long long, long long *) { return sn_fail(SN_ERR_UNSUPPORTED, "tiled kernel not built"); }
