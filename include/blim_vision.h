/*
 * blim_vision -- C ABI of the B200-native video feature extractor (part of libblim_b200.so).
 *
 * Scope: the step BEFORE the scoring path -- what the reference's extract.py:96-110 computes with
 * `model.encode_video_image(video, idx, return_video_feature=True)`:
 *     frames [n_frames, 3, S, S]  ->  UMT ViT encoder over clips of `frames_per_clip` frames (vision_tower_builder.py:564-577,
 *     329-348: Conv3d patch embedding, sinusoid position table, pre-norm blocks with LayerNorm / biased QKV / GELU MLP,
 *     final LayerNorm)  ->  ToMe bipartite token merging down to 16 tokens per frame (mm_projector_builder.py:6-130)
 *     ->  features [n_clips, 16 * frames_per_clip, C]  (the `.pth` files dataloader/base_dataset.py:26-31 loads).
 * Video decoding and the PIL resize / normalise of UMTImageProcessor stay on the host (extract.py:42-62).
 *
 * Conventions as in blim_b200.h: one extractor per process / device, work enqueued on the caller's stream, device
 * pointers owned by the caller, 0 = ok, blim_vision_last_error() for the message, dtype codes 0 = float32,
 * 1 = bfloat16, 2 = float16.
 */
#ifndef BLIM_VISION_H_
#define BLIM_VISION_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct blim_vision blim_vision;

/* Architecture of the UMT vision tower as build_vit() instantiates it (vision_tower_builder.py:506-523) plus the
 * token-merging target of ToMe16_mlp_hd64 (mm_projector_builder.py:146-154). */
typedef struct blim_vision_cfg {
  int32_t image_size;       /* 448 for VideoChat-Flash-Qwen2-7B_res448 */
  int32_t patch_size;       /* 16 */
  int32_t frames_per_clip;  /* mm_local_num_frames = 4 */
  int32_t hidden_size;      /* 1024, multiple of 64 */
  int32_t num_layers;       /* blocks actually run: encoder_depth + mm_vision_select_layer + 1 = 23 (vision_tower_builder.py:289) */
  int32_t num_heads;        /* 16; head_dim = hidden_size / num_heads must be 64 or 128 */
  int32_t mlp_hidden_size;  /* 4096 */
  int32_t tome_tokens_per_frame; /* 16: merge_tokens target = 16 * frames_per_clip (mm_projector_builder.py:147) */
  int32_t max_clips;        /* workspace: clips per blim_vision_extract call (0 = default 16) */
  float ln_eps;             /* block LayerNorms: 1e-6 (vision_tower_builder.py:371) */
  float final_ln_eps;       /* vision_layernorm: 1e-12 (vision_tower_builder.py:317) */
} blim_vision_cfg;

/* Replaces build_vision_tower + UMTVisionTower.load_model (vision_tower_builder.py:554-562, 609-618). */
int blim_vision_create(const blim_vision_cfg* cfg, int device, blim_vision** out);
void blim_vision_destroy(blim_vision* v);
const char* blim_vision_last_error(const blim_vision* v); /* v may be NULL: last create error */

/* Load one parameter by its reference state_dict key below `model.vision_tower.vision_tower.` (the prefix is optional):
 * "encoder.patch_embed.proj.weight" [C,3,1,P,P] / ".bias", "encoder.blocks.{i}.norm1|norm2.weight|bias",
 * "encoder.blocks.{i}.attn.qkv.weight" [3C,C], ".attn.q_bias", ".attn.v_bias" (the key bias is zero,
 * vision_tower_builder.py:101-103), ".attn.proj.weight|bias", ".mlp.fc1.weight|bias", ".mlp.fc2.weight|bias",
 * "encoder.vision_layernorm.weight|bias".  Blocks >= num_layers are accepted and ignored. */
int blim_vision_load_weight(blim_vision* v, const char* name, const void* dev_ptr, int dtype, const int64_t* shape, int ndim, void* stream);

/* Position table [frames_per_clip * (image_size/patch_size)^2, C] fp32 on the device: what
 * get_sinusoid_encoding_table / get_sinusoid_encoding_table2 return (vision_tower_builder.py:191-268); it is a
 * non-persistent attribute of the reference module, built on the host by blim_b200.vision.position_table. */
int blim_vision_set_pos_embed(blim_vision* v, const float* table_dev, int rows, void* stream);

/* UMTVisionTower.forward (vision_tower_builder.py:564-577): frames [n_frames, 3, S, S] (n_frames a multiple of
 * frames_per_clip, clip c = frames [c*fpc, (c+1)*fpc)) -> final-LayerNorm states feat_out fp32
 * [n_frames * (S/P)^2, C], token order (clip, frame, patch row, patch column). */
int blim_vision_encode(blim_vision* v, const void* frames_dev, int dtype, int n_frames, float* feat_out_dev, void* stream);

/* ToMe16_mlp_hd64.merge_tokens (mm_projector_builder.py:101-130) in fp32: x [b, p, C] -> out [b, target, C]; optional
 * debug outputs of the FIRST merging round: edge_idx [b, ceil(p/2)] (argsort of node_max, descending) and node_idx
 * [b, ceil(p/2)] (best partner of every even token), both int32. */
int blim_vision_merge_tokens(blim_vision* v, const float* x_dev, int b, int p, int target, float* out_dev, int32_t* edge_idx_out_dev,
                             int32_t* node_idx_out_dev, void* stream);

/* extract.py:100-106: encode + merge; out [n_clips, tome_tokens_per_frame * frames_per_clip, C] in out_dtype. */
int blim_vision_extract(blim_vision* v, const void* frames_dev, int dtype, int n_frames, void* out_dev, int out_dtype, void* stream);

/* Counters / timing for tools/extract_bench.py: kernels launched since creation, executed GEMM FLOPs; with profiling
 * enabled every launch family is bracketed by CUDA events on the launching stream.  read: index 0 tcgen05 GEMMs,
 * 1 attention, 2 LayerNorm, 3 patchify + position add, 4 token merging; synchronises the device and resets. */
int64_t blim_vision_kernel_launches(const blim_vision* v);
double blim_vision_gemm_flops(const blim_vision* v);
int blim_vision_profile(blim_vision* v, int enable);
int blim_vision_profile_read(blim_vision* v, int n, double* ms, int64_t* launches);

#ifdef __cplusplus
}
#endif
#endif /* BLIM_VISION_H_ */
