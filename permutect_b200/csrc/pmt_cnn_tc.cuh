// Declarations shared by the tensor-core haplotype-CNN forward (pmt_cnn_tc.cu) and the CNN backward that consumes the
// activations it saves (pmt_cnn_bwd.cu).
#pragma once
#include <cstring>

#include "pmt_host.h"
#include "pmt_tc_ptx.cuh"

namespace pmt {
namespace cnntc {

constexpr int EPI_WARPS = 16;
constexpr int THREADS = 32 * (EPI_WARPS + 1);
constexpr int MMA_WARP = EPI_WARPS;
constexpr int MAX_LAYERS = 10;
constexpr int MAX_CHUNKS = 4;
constexpr int CHUNK_COLS = 128;              // TMEM columns per chunk accumulator: [hi part N | lo part N]
constexpr int PLANE_ROWS = 344;
constexpr int PLANE_BYTES = PLANE_ROWS * 16;
constexpr int BUF_BYTES = 8 * PLANE_BYTES;   // one activation buffer: 32 channels
constexpr int C0 = 10;                       // one-hot channels: 2 haplotypes x 5 codes
constexpr int MAX_ITEMS = 24;                // (layer, chunk) work items per group
constexpr int MAX_TASKS = 8;                 // one-hot entries per lane in the im2col scatter

struct Layer {
  int first;      // im2col'd one-hot conv
  int taps;       // shifted-window convs: kernel size; first: number of input positions per row (ksize + dup)
  int ksteps;     // first: k-steps of 8 columns
  int N;          // output columns of the MMA (32, or 64 when dup)
  int dup, pool2; // fused MaxPool(2,1) after the first conv / MaxPool(2,2)
  int L_in, L_out, L_pool, L_next;
  int inv_L;      // ceil(65536 / L_in) + : row / L_in == (row * inv_L) >> 16 for rows < 512
  int act, to_global, out_ch;
  int img_off, img_bytes;
  // packing
  int op, in_ch, ksize, flat_len, scale_in, is_linear;
};

struct Item {
  int layer, chunk;
  int need;   // index of the last done_bar this item's MMAs must have observed (0 = im2col, 1 + i = epilogue of item i)
};

struct Plan {
  int n_layers, n_items, G, L0, image_bytes;
  int n_chunks[MAX_LAYERS];
  Layer layer[MAX_LAYERS];
  Item item[MAX_ITEMS];
};

// ---- training: what the SAVE variant of the forward leaves in global memory for the backward (pmt_cnn_bwd.cu) ----
// Per group of G variants and per layer l that feeds another layer: the stored activation a_l = what the forward keeps in
// its shared-memory planes (bias added, pooled, selu_u applied: the SELU scale lives in the consuming weights) as
// [32 channels][G * L_next] floats, and -- for layers with a fused max-pool of two -- one byte per (channel group of 8,
// position) whose bit i tells that the SECOND element of the window won for channel 8 cg + i: [4][G * L_next] bytes.
struct SaveLayout {
  int G, n_layers;
  int group_floats;             // floats per group
  int a_off[MAX_LAYERS];        // float offset of a_l inside a group (-1: not saved)
  int bits_off[MAX_LAYERS];     // float offset of the window bits (-1: no pool)
};

}  // namespace cnntc
}  // namespace pmt

// host (pmt_cnn_tc.cu)
bool pmt_build_cnn_tc_plan(const pmt::Plan& P, pmt::cnntc::Plan* out);
void pmt_cnn_save_layout(const pmt::cnntc::Plan& T, pmt::cnntc::SaveLayout* out);
int pmt_pack_cnn_tc_images(const pmt::Plan& P, const pmt::cnntc::Plan& T, const float* weights, unsigned char* image, cudaStream_t st);
// the forward over variants [v_first, v_first + n) of the batch with every layer's activations saved; `image` must hold
// the packed weight images (pmt_launch_cnn_tc with reuse_image = false packs them)
int pmt_launch_cnn_tc_save(const pmt::Plan& P, const pmt::cnntc::Plan& T, const float* weights, const PmtBatch* batch, int v_first, int n,
                           float* info_seq, const unsigned char* image, float* save, int n_sm, cudaStream_t st);
