// Packed frame layout shared by every decoder-side kernel.
//
// The reference keeps activations as zero-padded (B, C, T') tensors and masks them layer by layer
// (reference common.py:567, partialconv1d.py).  Padded frames never influence valid ones (SURVEY Appendix
// A-11), so on the GPU we pack only the valid frames of all utterances into ONE channels-last row axis:
//
//     | GAP zeros | utt 0 (len_0 rows) | GAP zeros | utt 1 | ... | GAP zeros | tail zeros up to rows_alloc |
//
// GAP = 16 rows >= the largest conv half-width (dilation 8 x 2 taps), so a shifted tap of one utterance can
// only ever land on zero rows, never on a neighbour.  Row metadata lets epilogues recover the position of a
// row inside its utterance (for the partial-conv renormalisation) without any host synchronisation.
//
// int32 plan buffer:  hdr[8] = {rows_used, B, Tmax, rows_alloc, 0,0,0,0}; row0[B]; len[B]; glen[B]; pad to 4;
//                     pos[rows_alloc]; rem[rows_alloc]; utt[rows_alloc]; tpos[rows_alloc]
// pos/rem describe VALID frames (-1 elsewhere); utt/tpos describe the geometric span of an utterance, which may be
// longer than its valid length when a layer must see the padded region too (plain, non-partial convs).
#pragma once
#include "common.cuh"

namespace rb {

constexpr int kGap = 16;
constexpr int kRowTile = 128;

__host__ __device__ inline int plan_rows_alloc(int B, int Tmax) { return round_up(B * (Tmax + kGap) + kGap, kRowTile); }
__host__ __device__ inline int plan_meta_off(int B) { return round_up(8 + 3 * B, 4); }
__host__ __device__ inline size_t plan_ints(int B, int Tmax) {
  return (size_t)plan_meta_off(B) + 4 * (size_t)plan_rows_alloc(B, Tmax);
}

struct PlanView {
  const int* base;
  int B, Tmax, rows_alloc;
  __host__ __device__ const int* hdr() const { return base; }
  __host__ __device__ const int* row0() const { return base + 8; }
  __host__ __device__ const int* len() const { return base + 8 + B; }
  __host__ __device__ const int* glen() const { return base + 8 + 2 * B; }
  __host__ __device__ const int* pos() const { return base + plan_meta_off(B); }
  __host__ __device__ const int* rem() const { return pos() + rows_alloc; }
  __host__ __device__ const int* utt() const { return rem() + rows_alloc; }
  __host__ __device__ const int* tpos() const { return utt() + rows_alloc; }
};

inline PlanView make_plan_view(const void* plan, int B, int Tmax) {
  PlanView v;
  v.base = reinterpret_cast<const int*>(plan);
  v.B = B;
  v.Tmax = Tmax;
  v.rows_alloc = plan_rows_alloc(B, Tmax);
  return v;
}

}  // namespace rb
