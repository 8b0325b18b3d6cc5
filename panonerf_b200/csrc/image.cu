// Render driver and validation outputs: the per-ray results of the hot path are
// scattered straight into the [C,H,W] images the systems return (systems/panonerf_system.py:171-189 builds them
// with cat + view + permute), image metrics are reduced on the device (utils/metrics.py:210-237, 318-326) and the
// OpenEXR / PNG payloads are laid out on the device (utils/io_exr.py:30-47, utils/vis.py:25-41) so that validation
// needs one device-to-host copy per file and no host-side pixel loops.  All HBM-bound element-wise maps.
#include "common.cuh"

namespace pnb {

constexpr int kMaxImages = 10;
struct ChwArgs {
  const float* src[kMaxImages];  // per-ray results [R, ch[i]] (row-major)
  int ch[kMaxImages];            // channels of image i
  int coff[kMaxImages];          // first output plane of image i
  int n_img, total_ch;
};

// out[(coff_i + c) * HW + pix0 + r] = src_i[r * ch_i + c]
__global__ void pack_chw_kernel(ChwArgs a, long long R, long long HW, long long pix0, float* __restrict__ out) {
  const long long total = R * a.total_ch;
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const int plane = (int)(idx / R);
    const long long r = idx - (long long)plane * R;
    int i = 0;
#pragma unroll
    for (int k = 1; k < kMaxImages; ++k)
      if (k < a.n_img && plane >= a.coff[k]) i = k;
    const int c = plane - a.coff[i];
    out[(long long)plane * HW + pix0 + r] = a.src[i][r * a.ch[i] + c];
  }
}

// (pred - gt)^2 [* row weight]  ->  per-element terms (summed in a fixed order by pnb_sum)
__global__ void sqerr_kernel(long long n, int W, long long HW, const float* __restrict__ pred,
                             const float* __restrict__ gt, const float* __restrict__ row_w, float* __restrict__ out) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const float d = pred[i] - gt[i];
    float v = d * d;
    if (row_w != nullptr) v = v * row_w[(i % HW) / W];
    out[i] = v;
  }
}

// Scan-line blocks of an uncompressed OpenEXR file with FLOAT channels B, G, R (alphabetical, as the format requires):
// per line [int32 y][int32 bytes][B row][G row][R row].  C == 1 replicates the channel (utils/io_exr.py:43-44).
__global__ void exr_pack_kernel(int H, int W, int C, const float* __restrict__ chw, uint32_t* __restrict__ out) {
  const long long line_words = 2 + 3ll * W;
  const long long total = (long long)H * line_words;
  const long long HW = (long long)H * W;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int y = (int)(i / line_words);
    const long long k = i - (long long)y * line_words;
    uint32_t v;
    if (k == 0) v = (uint32_t)y;
    else if (k == 1) v = (uint32_t)(3 * W * 4);
    else {
      const int plane = (int)((k - 2) / W);          // 0: B, 1: G, 2: R
      const int x = (int)((k - 2) - (long long)plane * W);
      const int c = C == 1 ? 0 : 2 - plane;
      v = __float_as_uint(chw[(long long)c * HW + (long long)y * W + x]);
    }
    out[i] = v;
  }
}

// Filtered PNG scan lines (filter type 0) of an 8-bit RGB image: (image * 255) truncated like numpy's astype(uint8)
// on [0, 1] (utils/vis.py:35); values outside are clamped.  C == 1 replicates the channel (vis.py:31-32).
__global__ void png_pack_kernel(int H, int W, int C, const float* __restrict__ chw, uint8_t* __restrict__ out) {
  const long long line = 1 + 3ll * W;
  const long long total = (long long)H * line;
  const long long HW = (long long)H * W;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int y = (int)(i / line);
    const long long k = i - (long long)y * line;
    uint8_t v = 0;
    if (k > 0) {
      const int x = (int)((k - 1) / 3), c = (int)((k - 1) - 3ll * x);
      float f = chw[(long long)(C == 1 ? 0 : c) * HW + (long long)y * W + x] * 255.f;
      f = fminf(fmaxf(f, 0.f), 255.f);
      v = (uint8_t)(int)f;  // truncation toward zero
    }
    out[i] = v;
  }
}

}  // namespace pnb

using namespace pnb;

extern "C" int pnb_pack_chw(long long R, long long HW, long long pix0, int n_img, const void* const* src_host,
                            const int* channels_host, float* out, void* stream) {
  PNB_REQUIRE(R >= 0 && HW > 0 && pix0 >= 0 && pix0 + R <= HW, "pack_chw: ray range outside the image");
  PNB_REQUIRE(n_img >= 1 && n_img <= kMaxImages && src_host && channels_host && out, "pack_chw: bad arguments");
  if (R == 0) return 0;
  ChwArgs a{};
  a.n_img = n_img;
  int off = 0;
  for (int i = 0; i < n_img; ++i) {
    PNB_REQUIRE(src_host[i] != nullptr && channels_host[i] >= 1, "pack_chw: null image or bad channel count");
    a.src[i] = reinterpret_cast<const float*>(src_host[i]);
    a.ch[i] = channels_host[i];
    a.coff[i] = off;
    off += channels_host[i];
  }
  a.total_ch = off;
  pack_chw_kernel<<<grid_for(R * off, 256), 256, 0, as_stream(stream)>>>(a, R, HW, pix0, out);
  return finish("pack_chw");
}

extern "C" int pnb_image_sqerr(long long n, int W, long long HW, const float* pred, const float* gt,
                               const float* row_weights, float* out, void* stream) {
  PNB_REQUIRE(n >= 0 && W > 0 && HW > 0 && pred && gt && out, "image_sqerr: bad arguments");
  if (n == 0) return 0;
  sqerr_kernel<<<grid_for(n, 256), 256, 0, as_stream(stream)>>>(n, W, HW, pred, gt, row_weights, out);
  return finish("image_sqerr");
}

extern "C" long long pnb_exr_payload_bytes(int H, int W) { return (long long)H * (8 + 12ll * W); }

extern "C" int pnb_exr_pack(int H, int W, int C, const float* chw, void* out, void* stream) {
  PNB_REQUIRE(H > 0 && W > 0 && (C == 1 || C == 3) && chw && out, "exr_pack: need a 1- or 3-channel image");
  exr_pack_kernel<<<grid_for((long long)H * (2 + 3ll * W), 256), 256, 0, as_stream(stream)>>>(
      H, W, C, chw, reinterpret_cast<uint32_t*>(out));
  return finish("exr_pack");
}

extern "C" long long pnb_png_payload_bytes(int H, int W) { return (long long)H * (1 + 3ll * W); }

extern "C" int pnb_png_pack(int H, int W, int C, const float* chw, void* out, void* stream) {
  PNB_REQUIRE(H > 0 && W > 0 && (C == 1 || C == 3) && chw && out, "png_pack: need a 1- or 3-channel image");
  png_pack_kernel<<<grid_for((long long)H * (1 + 3ll * W), 256), 256, 0, as_stream(stream)>>>(
      H, W, C, chw, reinterpret_cast<uint8_t*>(out));
  return finish("png_pack");
}
