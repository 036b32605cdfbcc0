// Kernel parameter blocks shared between encoder.cu and the C-ABI layer.
#pragma once
#include <cstdint>

namespace ldot {

struct EmbedImageParams {
  const float* lin;       // [B*R, H] fp32: img_linear(feat) + bias (tcgen05 GEMM output)
  const float* box;       // [B*R, 7] fp32 (x1, y1, x2, y2, w, h, w*h)
  const float* img_g; const float* img_b;   // img_layer_norm
  const float* pos_w;     // [H, 7]
  const float* pos_bias;  // [H]
  const float* pos_g; const float* pos_b;   // pos_layer_norm
  const float* type1;     // [H] token_type_embeddings[1]
  const float* ln_g; const float* ln_b;     // img_embeddings.LayerNorm
  uint16_t* out;          // [B, out_seq, H]; region r of image b -> row b * out_seq + row_offset + r
  int B, R, out_seq, row_offset;
};

}  // namespace ldot
