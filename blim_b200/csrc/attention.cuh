// Sequence description shared by the host-side scheduler and the tcgen05 attention kernels (attention_tc.cuh), plus the
// cp.async helpers they use.
//
// The attention replaces Qwen2SdpaAttention's softmax(QK^T/sqrt(d) + mask)V (reference: modeling_qwen2_flash.py:685-709)
// and repeat_kv (modeling_qwen2_flash.py:192-201, never materialised: the G query heads of a KV group are stacked along
// the row dimension of one tile).  Every sequence sees
//   segment A: a_len keys of a previously prefilled prefix (all visible) -- the video prefix shared by all captions of a
//              video (VTG), the text prefix shared by all candidate videos of a text (TVG), or the CPN-visible header;
//   segment B: its own q_len tokens, key j visible to query i iff j <= i and key_valid[j]
// which is exactly "causal AND key-valid" of the reference's 4-D additive mask (modeling_qwen2_flash.py:1019-1040) once
// the invisible tokens are dropped from the layout.  Rotary positions were applied by the QKV epilogue, so position
// gaps (CPN) need nothing here.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <vector>

namespace blim {

struct AttnSeq {
  int q_start;  // first token of the sequence in this run (row of Q / O)
  int q_len;
  int a_start;  // first prefix key row in k_a / v_a
  int a_len;
  int b_start;  // first own key row in k_b / v_b (its q_len tokens are consecutive rows)
  int unit;     // scheduling unit (video / text) the sequence belongs to: attention tiles never stack sequences of different
                // units, so a unit's numbers do not depend on which other units share its run (multi-GPU sharding, batching)
};

// 16-byte async copy global -> shared; src_bytes = 0 zero-fills the destination.
__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc, int src_bytes) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(static_cast<uint32_t>(__cvta_generic_to_shared(smem_dst))), "l"(gsrc),
               "r"(src_bytes)
               : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

}  // namespace blim
