// Internal: the single-pass-per-radix-level decode for large top_n (decode_large.cu), called by ssdsb_decode.
#pragma once
#include <cuda_runtime.h>

#include "common.cuh"

namespace ssdsb {

int decode_large_max_k();                                   // largest top_n the large path sorts in shared memory
size_t decode_large_workspace_bytes(const ssdsb_level* levels, int n_levels, int B, int top_n);
int decode_large(const ssdsb_level* levels, int n_levels, int B, float threshold, int top_n, int rescore,
                 float* d_scores, float* d_boxes, float* d_classes, int32_t* d_index, void* d_workspace,
                 size_t workspace_bytes, cudaStream_t st);

// Sorted top-K keys (make_key(score, index), descending) of every row of d_scores [B, N] among scores >= min_score:
// d_keys [B, K] (zero padded), d_count [B] = number of valid keys per row.  Used by nms for long candidate rows.
size_t topk_rows_workspace_bytes(int B, int N, int K);
int topk_rows(const float* d_scores, int B, int N, float min_score, int K, unsigned long long* d_keys, int* d_count,
              void* d_workspace, size_t workspace_bytes, cudaStream_t st);

}  // namespace ssdsb
