// nm_match_tc.cu -- tensor-core (tcgen05/TMEM) candidate search for the matcher.
// Placeholder while the kernel is brought up: reports "unavailable", so nm_match_*
// use the exact fp32 engine (nm_match.cu).
#include "nm_match.cuh"

bool nm_match_tc_available() { return false; }

int nm_match_scan_tc(const float*, int, const float*, int, int, float4*, cudaStream_t)
{
    return NM_ERR_UNSUPPORTED;
}
