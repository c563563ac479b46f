// nm_match.cuh -- internal interface of the matcher kernels.
#pragma once
#include "nm_common.cuh"

// The reference's min2 start value: the int literal 0x7f800000 converted to float
// (gpu/kernels/match.cu:91), NOT +inf.
#define NM_MIN2_INIT 2139095040.0f

// Exact fp32 SIMT scan: per-row true top-2 records rec4[a] = (d1, bits(i1+index_offset), d2, 0)
// over B rows [0,nB); optionally also writes the distance matrix D[a*d_sa + b*d_sb].
// A(a,k) = A[a*a_sa + k*a_sk].
int nm_match_scan_exact(const float* A, long long a_sa, long long a_sk, int nA, const float* B, int nB,
                        int dim, int index_offset, float4* rec4, float* D, long long d_sa, long long d_sb,
                        cudaStream_t stream);

// The same exact scan restricted to the query rows row_list[0 .. *row_count) (both on the device).
int nm_match_scan_exact_rows(const float* A, int nA, const float* B, int nB, int dim, int index_offset,
                             const int* row_list, const int* row_count, float4* rec4, cudaStream_t stream);

// Tensor-core candidate search + exact re-rank (nm_match_tc.cu).  Same record contract.
int nm_match_scan_tc(const float* A, int nA, const float* B, int nB, int index_offset, float4* rec4,
                     cudaStream_t stream);
bool nm_match_tc_available();

// shard_stride = rows between the record arrays of consecutive shards (0: nA, i.e. densely packed)
int nm_match_finalize(const float4* recs, int n_shards, int nA, float ambiguity, int* match_io,
                      cudaStream_t stream, long long shard_stride = 0);
