/* deflow_b200 -- C ABI of the B200-native DeFlow hot path.
 *
 * Plain pointers and sizes only (no torch types).  Every pointer is a DEVICE pointer unless the
 * comment says HOST.  Every function is asynchronous on the given stream (a cudaStream_t passed
 * as void*), returns 0 on success and a DFB_ERR_* code otherwise; dfb_last_error() returns the
 * message of the last failure on the calling thread.  Nothing here allocates device memory.
 *
 * REF = /root/reference/OpenSceneFlow.  Each entry point names the reference interface it
 * replaces; INTEGRATION.md shows the binding a reference maintainer would add.
 */
#ifndef DEFLOW_B200_H
#define DEFLOW_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ------------------------------------------------------------------ library */
const char* dfb_last_error(void);
int dfb_version(void);
/* Number of kernels this library has launched on the calling process so far (bench.py's
 * "gpu_launches" is the difference across the timed region). */
long long dfb_launch_count(void);

/* ------------------------------------------------------------------ mmcv._ext drop-ins
 * REF/assets/cuda/mmcv/pybind.cpp:32-50 -- the four functions `mmcv._ext` exports.
 * hard_voxelize_forward is not on the DeFlow path (HardVoxelizer is unused) and is not provided.
 */

/* grid = round((max-min)/voxel) in fp32 (REF/assets/cuda/mmcv/voxelization_cuda.cu:269-271).
 * voxel_size[3], range[6], grid_xyz[3] are HOST pointers. */
int dfb_grid_size(const float* voxel_size, const float* range, int* grid_xyz);

/* dynamic_voxelize_forward(points, voxel_size, coors_range, coors, NDim=3)
 * (REF/assets/cuda/mmcv/voxelization.cpp:62-74, voxelization_cuda_kernel.cuh:13-50).
 * points f32[n, num_features>=3]; coors i32[n,3] pre-zeroed by the caller, written in (z,y,x)
 * order with the reference's partial -1 pattern.  voxel_size/range: HOST fp32. */
int dfb_dynamic_voxelize_forward(const float* points, int n, int num_features, const float* voxel_size,
                                 const float* range, int* coors, void* stream);

/* dynamic_point_to_voxel_forward(feats, coors, reduce_type)
 * (REF/assets/cuda/mmcv/scatter_points.cpp:36-41, scatter_points_cuda.cu:9-66) in two calls,
 * because the number of pillars M is only known on the device:
 *   1. dfb_scatter_index   builds point2voxel_map[n] (-1 for rows with a negative component),
 *      voxel_coors[<=n,3] sorted lexicographically, voxel_points_count[<=n], the CSR arrays
 *      pil_start[<=n+1] / sorted_pt[n], and counts6 (device i32[6]: [0] = number of valid points,
 *      [1] = number of voxels M; the rest is scratch).  extent[3] (HOST) is an
 *      exclusive upper bound of the coordinates per column (z,y,x); coordinates >= extent are a
 *      caller error.  Workspace: bitmap and word_rank of ceil(prod(extent)/32) words each,
 *      slot[n], blk[ceil(n/1024)].
 *   2. dfb_scatter_reduce  reduces feats f32[n,c] into voxel_feats f32[M,c];
 *      reduce_type 0 = sum, 1 = mean, 2 = max (REF reduce_t).
 */
int dfb_scatter_index(const int* coors, int n, const int* extent_zyx, unsigned* bitmap, int* word_rank, int* blk,
                      int* slot, int* point2voxel_map, int* voxel_coors, int* voxel_points_count, int* pil_start,
                      int* sorted_pt, int* counts6, void* stream);
int dfb_scatter_reduce(const float* feats, int n, int c, const int* pil_start, const int* sorted_pt,
                       const int* num_voxels, int max_voxels, int reduce_type, float* voxel_feats, void* stream);

/* dynamic_point_to_voxel_backward(grad_feats, grad_reduced_feats, feats, reduced_feats, coors_idx,
 * reduce_count, reduce_type) (REF/assets/cuda/mmcv/scatter_points.cpp:43-53,
 * scatter_points_cuda.cu:68-132).  grad_feats f32[n,c] is fully written (zeros where map == -1).
 * workspace_mc: i32[m*c], only needed for reduce_type 2 (max trace-back), else may be NULL. */
int dfb_dynamic_point_to_voxel_backward(float* grad_feats, const float* grad_reduced_feats, const float* feats,
                                        const float* reduced_feats, const int* point2voxel_map,
                                        const int* voxel_points_count, int n, int m, int c, int reduce_type,
                                        int* workspace_mc, void* stream);

/* ------------------------------------------------------------------ batched pillar index
 * Replaces DynamicVoxelizer.forward + the unique_dim half of DynamicScatter for all frames of a
 * step at once (REF/src/models/basic/encoder.py:567-600; scatter_points_cuda.cu:24-37).
 * Outputs are flat, frame after frame, valid points in their original order:
 *   counts   i32[4F+2] = n_valid[F] | n_pillars[F] | pt_off[F+1] | pil_off[F+1]
 *   pt_xyz   f32[cap,3]  pt_coor i32[cap,3] (z,y,x)  pt_idx i64[cap]  pt_offs f32[cap,3]
 *   pt_pillar i32[cap]   global pillar id (reference point2voxel_map = pt_pillar - pil_off[f])
 *   pil_cnt i32[pil_cap] pil_coor i32[pil_cap,3] pil_pix i32[pil_cap] (= f*H*W + y*W + x)
 *   pil_start i32[pil_cap+1], sorted_pt i32[cap]: CSR list of the points of every pillar
 *   csr_rec  f32[cap,4]: (x, y, z, pillar id) in CSR order -- what the feature-net passes stream through
 * with cap = F*Nmax.  Workspace sizes come from dfb_index_workspace().
 */
typedef struct {
  int F, Nmax, pt_stride;
  long long pil_cap;
  float voxel_size[3];
  float range[6];
  const float* pts;
  int* keys;         /* unused since r01 (the compaction pass recomputes voxels); may be NULL */
  unsigned* bitmap;  /* [F*words] */
  int* word_rank;    /* [F*words] */
  int* blk_cnt;      /* [F*blocks] */
  int* pt_slot;      /* [F*Nmax] */
  int* counts;
  float* pt_xyz;
  int* pt_coor;
  long long* pt_idx;
  float* pt_offs;
  int* pt_pillar;
  int* pil_cnt;
  int* pil_coor;
  int* pil_pix;
  int* pil_start;
  int* sorted_pt;
  float* csr_rec;    /* out [F*Nmax,4]: (x, y, z, bits of the global pillar id) of every valid point in CSR order */
  int* scan_ws;      /* workspace, dfb_index_scan_workspace() int32 elements */
  unsigned* tickets; /* workspace [2], zeroed inside */
  void* zero_base;   /* optional: when bitmap | pil_cnt | blk_cnt | tickets are carved from ONE allocation, its base ... */
  long long zero_bytes; /* ... and size, so that a single memset clears them (NULL: four memsets) */
  unsigned char* occ;   /* optional occupancy BYTE map [F * 32 * bitmap words] (part of zero_base, or zeroed inside): points mark
                           their cell with a plain byte store instead of an atomicOr on the bitmap; the scan packs it to bits */
} dfb_index_args;

int dfb_index_workspace(int F, int Nmax, const float* voxel_size, const float* range,
                        long long* bitmap_words_per_frame, long long* blocks_per_frame);
long long dfb_index_scan_workspace(int F, long long bitmap_words_per_frame, long long pil_cap);
int dfb_pillar_index(const dfb_index_args* args, void* stream);

/* ------------------------------------------------------------------ ego-motion compensation
 * cal_pose0to1 + warp (REF/src/models/basic/__init__.py:4-15, REF/src/models/deflow.py:60-77).
 * pose0/pose1 f32[B,4,4]; ego (optional, may be NULL) f32[B,4,4] overrides the pose product.
 * pc0 f32[B,Nmax,3] -> pose_flow (same shape) and pc0_warped, whose sample b starts at element
 * b * warped_stride_b (so it can be written straight into the [2B,Nmax',3] buffer the pillar index
 * consumes); NaN rows stay NaN.  pose_0to1 (optional out) f32[B,4,4]. */
int dfb_ego_warp(const float* pose0, const float* pose1, const float* ego, const float* pc0, int B, int Nmax,
                 float* pc0_warped, long long warped_stride_b, float* pose_flow, float* pose_0to1, void* stream);

/* ------------------------------------------------------------------ fused pillar feature net
 * DynamicPillarFeatureNet.forward + PointPillarsScatter (REF/src/models/basic/encoder.py:430-475,
 * 126-147) for every frame at once, never materialising the [N,9] / [N,32] point tensors:
 * decorate(9) -> Linear(9,32) -> BatchNorm1d(eps 1e-3) -> ReLU -> pillar mean -> NHWC image.
 * BatchNorm statistics are per frame, exactly like the reference's one-call-per-sample loop. */
typedef struct {
  int F, H, W;            /* frames, image rows (grid_y), image cols (grid_x) */
  int training;           /* 1: batch statistics per frame; 0: running statistics */
  float voxel_size[3];
  float center_off[3];    /* v/2 + range_min, computed in double by the caller (encoder.py:257-259) */
  float eps, momentum;
  const int* counts;
  const float* pt_xyz;
  const int* pt_coor;
  const int* pt_pillar;
  const int* pil_cnt;
  const int* pil_coor;
  const int* pil_pix;
  const int* pil_start;
  const int* sorted_pt;
  const float* weight;    /* [32,9] */
  const float* gamma;     /* [32] */
  const float* beta;      /* [32] */
  float* running_mean;    /* [32] updated in place when training (2B sequential updates) */
  float* running_var;     /* [32] */
  float* pil_mean;        /* out [pil_cap,3]: cluster_scatter result */
  double* stats;          /* workspace/out [F,64]: per-frame moments of the decorated features: S1[9] | S2[45] (upper triangle) */
  float* bn_params;       /* out [F,4,32]: a = gamma*rstd, b = beta - mean*a, mean, rstd (saved for backward) */
  float* pil_feats;       /* out [pil_cap,32] fp32: pfn_scatter result (voxel_feats); NULL = not wanted */
  void* image;            /* out [F,H,W,32] NHWC, zero-filled here; bf16 if image_bf16 else f32 */
  int image_bf16;
  long long pil_cap;      /* rows of pil_feats / pil_mean */
  const float* csr_rec;   /* [cap,4] from dfb_pillar_index: (x, y, z, pillar id) in CSR order */
  unsigned* pt_mask;      /* out [cap]: the 32 ReLU decisions of every point (CSR order), read by the backward */
  float* partials;        /* workspace [(cap/32 + 1), 2, 32] fp32: partial sums of pillars that straddle 32-point groups */
  float* pil_hdr;         /* out [pil_cap,12]: per pillar (first CSR position, count, pixel, frame | mean xyz, centre x | centre yz)
                             written by the forward, read by every later pass and the backward */
  void* image_ready_event;/* optional cudaEvent_t: the caller zero-fills `image` on another stream; the first kernel that
                             writes the image waits for this event and the call does not memset the image itself */
  float range_min[3];     /* point_cloud_range[0:3] (x, y, z minimum), fp32: the backward recomputes a point's pixel from it */
  /* SyncBatchNorm (REF/conf/config.yaml:23 sync_bn, REF/train.py:128): the call can run in two phases around the caller's
   * all-reduce.  phase 0 = everything (default); 1 = up to the per-frame moments in `stats`; 2 = from the BatchNorm
   * finalisation on, normalising with sync_stats [F,64] (the moments summed over all ranks) and sync_counts [F] (pooled
   * point counts) when given.  `stats` keeps this rank's moments (the backward needs them). */
  int phase;
  const double* sync_stats;
  const int* sync_counts;
} dfb_pfn_args;
int dfb_pfn_forward(const dfb_pfn_args* args, void* stream);
/* Dense zero fill (the PointPillarsScatter canvas) with blocks_per_sm small blocks per SM, meant to run on a second
 * stream under other kernels; ptr 16-byte aligned. */
int dfb_zero_fill(void* ptr, long long bytes, int blocks_per_sm, void* stream);
/* Sparse clear of a pseudo-image that is reused from call to call: zero the row_bytes-byte rows pix[0 .. n) with
 * n = min(counts[count_index], cap) -- the pillars the PREVIOUS forward wrote (its pil_pix / counts) -- instead of
 * zero-filling the whole dense canvas again (PointPillarsScatter's zeros() + index_put, encoder.py:135-141). */
int dfb_clear_rows(void* image, int row_bytes, const int* pix, const int* counts, int count_index, long long cap, void* stream);

typedef struct {
  dfb_pfn_args fwd;         /* same buffers as the forward call */
  const void* grad_image;   /* [F,H,W,32] NHWC (dtype = fwd.image_bf16 ? bf16 : f32) */
  float* grad_weight;       /* [32,9]  accumulated (+=) */
  float* grad_gamma;        /* [32]    accumulated (+=) */
  float* grad_beta;         /* [32]    accumulated (+=) */
  double* bwd_stats;        /* workspace [F,32,10] doubles: per frame and channel A1 | T[9] (zeroed inside) */
  double* grad_accum;       /* unused (kept for ABI stability) */
  int phase;                /* 0 = everything; 1 = point pass only (bwd_stats); 2 = finalisation only */
  const double* sync_bwd_stats; /* phase 2 under SyncBatchNorm: bwd_stats summed over all ranks (fwd.sync_counts = pooled counts) */
} dfb_pfn_bwd_args;
int dfb_pfn_backward(const dfb_pfn_bwd_args* args, void* stream);

/* ------------------------------------------------------------------ decoder gather
 * The advanced-index gather of ConvGRUDecoder / LinearDecoder (REF/src/models/basic/decoder.py:
 * 215-225): h0[p] = [img0[y,x,0:32], img1[y,x,0:32], unet[y,x,0:64]] for every pc0 point.
 * img: [2B,H,W,32] NHWC (frames 0..B-1 = pc0, B..2B-1 = pc1), unet: [B,H,W,64] NHWC.
 * Backward is a segment sum over the CSR pillar lists (no atomics) into dense NHWC gradients. */
int dfb_decoder_gather(const void* img, const void* unet, int in_bf16, int B, int H, int W, const int* counts,
                       int F, const int* pt_pillar, const int* pil_pix, void* h0, int out_bf16, int n_cap,
                       void* stream);
int dfb_decoder_gather_backward(const void* grad_h0, int grad_bf16, int B, int H, int W, const int* counts, int F,
                                const int* pil_pix, const int* pil_start, const int* sorted_pt, void* grad_img,
                                void* grad_unet, int out_bf16, int pil_cap, void* stream);
/* The same with the image part deferred: the 64 image channels of every pc0 pillar's gradient sum go to the compact buffer
 * img_rows f32[pil_cap][64] instead of a dense zero-filled image gradient (grad_unet stays dense); dfb_gather_img_rows_add adds
 * them into the image gradient [2B,H,W,32] in place once the other consumers of the pseudo-image have produced theirs:
 * no dense zero-fill, no dense addition.  unet_colsum: NULL, or f32[64] += per-channel sums of grad_unet over all pixels (the
 * bias gradient of the convolution that produced the UNet output, REF/src/models/basic/unet.py:100-103) from the pillar sums. */
int dfb_decoder_gather_backward_rows(const void* grad_h0, int grad_bf16, int B, int H, int W, const int* counts, int F,
                                     const int* pil_pix, const int* pil_start, const int* sorted_pt, float* img_rows,
                                     void* grad_unet, int out_bf16, int pil_cap, float* unet_colsum, void* stream);
int dfb_gather_img_rows_add(const float* img_rows, int B, int H, int W, const int* counts, int F, const int* pil_pix,
                            void* grad_img, int out_bf16, int pil_cap, void* stream);
/* out[0:half) = a0 + b0, out[half:2*half) = a1 + b1 (elementwise, bf16 or fp32; b0 = b1 = NULL: plain concatenation): the
 * gradient of the pseudo-image [2B,H,W,32] from the two consumers of each frame half in one pass. */
int dfb_add_cat2(const void* a0, const void* b0, const void* a1, const void* b1, long long bytes_per_half, int bf16, void* out,
                 void* stream);

/* ------------------------------------------------------------------ losses
 * deflowLoss / ff3dLoss with the trainer's gt construction fused in
 * (REF/src/lossfuncs.py:102-125, 148-157; REF/src/trainer.py:120-142):
 *   gt[p] = flow_gt[b, idx[p]] - pose_flow[b, idx[p]];  loss = SUM over samples.
 * kind 0 = deflowLoss, 1 = ff3dLoss.  est f32[n,3] flat over the B pc0 frames.
 * bucket_ws: double[B*8] workspace; loss: f32[1]; grad_est f32[n,3] (d loss / d est, may be NULL). */
int dfb_flow_loss(int kind, const float* est, const float* flow_gt, const float* pose_flow,
                  const unsigned char* classes, const long long* pt_idx, const int* counts, int F, int B,
                  int Nmax, double* bucket_ws, float* loss, float* grad_est, int n_cap, void* stream);

/* ------------------------------------------------------------------ UNet convolutions (tensor cores)
 * Every nn.Conv2d of FastFlow3DUNet (REF/src/models/basic/unet.py:49-68; ConvWithNorms in
 * REF/src/models/basic/__init__.py:61-79), as a tcgen05 implicit GEMM on NHWC bf16 tensors.
 *   dfb_conv_pack_weights: torch-layout fp32 weights [cout,cin,k,k] -> bf16 GEMM operands
 *       w_fwd [cout][tap*cin + ci] and w_dgrad [cin][tap*cout + co] (either may be NULL).
 *   dfb_conv2d mode 0 (forward): y[n,Ho,Wo,cout] = conv(cat(x[0..n_src-1], channel axis)) + bias, optional
 *       per-channel sum / sum-of-squares of y accumulated (+=) into stats[2][cout] (BatchNorm batch statistics).
 *   dfb_conv2d mode 1 (data gradient): x[0] = grad_y [n,Ho,Wo,cout] -> y = grad_x [n,H,W,cin[0]] for the input
 *       channel slice [cin_off, cin_off + cin[0]) of a layer with cin_total input channels.
 *   dfb_conv2d_wgrad: grad_w (+)= sum over pixels, written in torch layout fp32 [cout,cin_total,k,k].
 * H, W are always the spatial size of the convolution INPUT; padding = ksize/2. */
typedef struct {
  int mode;              /* 0 forward, 1 data gradient */
  int n, H, W;           /* batch, input height, input width */
  int ksize, stride;     /* 1 or 3; 1 or 2 */
  int n_src;             /* forward: 1 or 2 channel-concatenated sources */
  const void* x[2];      /* bf16 NHWC sources */
  int cin[2];            /* channels per source (multiples of 32) */
  int cin_total, cin_off;/* dgrad only: the layer's total input channels and this slice's offset */
  int cout;              /* 32, 64, 128 or 256 (forward N tile = cout) */
  const void* w;         /* packed bf16 weights: w_fwd (mode 0) or w_dgrad (mode 1) */
  const float* bias;     /* [cout] fp32 or NULL (mode 0) */
  void* y;               /* output NHWC, bf16 (or fp32 if y_fp32) */
  int y_fp32;
  double* stats;         /* [2][cout] or NULL */
  const void* x_lo[2];   /* split-precision mode: the lo halves of the sources (x[] are the hi halves) */
  int split3;            /* 1: operands are (hi, lo) bf16 pairs, weights packed with split3, K loop = hi*hi + hi*lo + lo*hi */
  int stats_sum_only;    /* 1: only stats[0][*] (the per-channel sums) is wanted, e.g. a bias gradient from a data-gradient launch */
  void* y2;              /* mode 1, 1x1 stride 1, bf16 only: second output.  The launch then covers the input channels
                            [cin_off, cin_off + cin[0] + cin2): the first cin[0] go to y [n,H,W,cin[0]], the next cin2 to
                            y2 [n,H,W,cin2] -- the two sources of a concatenated input get their gradients from ONE pass
                            over grad_y */
  int cin2;
  float* grad_bias;      /* dfb_conv2d_wgrad, 1x1 stride 1 with an odd number of 64-channel input groups only: [cout] fp32
                            += sum over pixels of grad_y (the layer's bias gradient, from otherwise idle accumulator rows
                            of the same launch); NULL = not wanted */
} dfb_conv_args;
int dfb_conv_pack_weights(const float* w, int cout, int cin, int ksize, int split3, void* w_fwd, void* w_dgrad,
                          void* stream);
/* Every convolution weight of a model in ONE launch (the training step re-packs after each optimizer update:
 * the fp32 masters in torch layout are the state_dict contract, REF/src/models/basic/unet.py:49-68).  table_dev: device
 * array of n_weights (<= 64) descriptors, `first` = running element offset (exclusive prefix sum of cout*cin*k*k);
 * total_elems = sum of all element counts. */
typedef struct dfb_pack_desc {
  const float* w;        /* [cout,cin,k,k] fp32 */
  void* w_fwd;           /* bf16 [cout][(split3?2:1)*k*k*cin] or NULL */
  void* w_dgrad;         /* bf16 [cin][(split3?2:1)*k*k*cout] or NULL */
  long long first;
  int cout, cin, ksize, pad_;
} dfb_pack_desc;
int dfb_conv_pack_weights_multi(const dfb_pack_desc* table_dev, int n_weights, long long total_elems, int split3,
                                void* stream);
/* The way back for the gradients: every weight gradient of a model from its [tap][cout][cin] accumulator (dfb_conv2d_wgrad
 * with grad_w == NULL leaves it there; accumulate bit 1 = do not zero the accumulator first, so the two calls of a shared
 * encoder weight and the three products of the split-precision mode sum in place) to the torch layout, in ONE launch. */
typedef struct dfb_unpack_desc {
  const float* wacc;     /* [k*k][cout][cin] fp32 */
  float* grad;           /* [cout,cin,k,k] fp32 */
  long long first;
  int cout, cin, ksize, accumulate;   /* accumulate: grad += instead of grad = */
} dfb_unpack_desc;
int dfb_wgrad_unpack_multi(const dfb_unpack_desc* table_dev, int n_weights, long long total_elems, void* stream);
/* fp32 x[n] -> hi = bf16(x), lo = bf16(x - hi): the operand pairs of the split-precision ("bf16x3") parity mode */
int dfb_split_bf16x2(const float* x, long long n, void* hi, void* lo, void* stream);
int dfb_conv2d(const dfb_conv_args* args, void* stream);
/* args as for the forward (x sources, geometry) with args->y = grad_y (bf16, input).  wacc: fp32 workspace
 * [k*k][cout][cin_total] (zeroed inside); grad_w: fp32 torch layout, overwritten or (accumulate) added to. */
int dfb_conv2d_wgrad(const dfb_conv_args* args, float* wacc, float* grad_w, int accumulate, void* stream);

/* ------------------------------------------------------------------ UNet HBM-bound passes (NHWC bf16)
 * BatchNorm2d(eps 1e-5, momentum 0.1) + exact GELU of ConvWithNorms (REF/src/models/basic/__init__.py:61-79) and
 * the bilinear x2 upsample (REF/src/models/basic/unet.py:8-18).
 *   dfb_bn2d_finalize: stats[2][C] (sum, sum of squares over `count` elements, from dfb_conv2d) -> bn[4][C] =
 *       a = gamma*rstd | b = beta - mean*a | mean | rstd; training updates running_mean / running_var in place.
 *   dfb_bn_gelu_apply: y = GELU(a*x + b).
 *   dfb_bn_gelu_backward: gx = dL/dx from gy = dL/dy; g_gamma / g_beta / g_bias (+=); red: double[2*C] workspace.
 *   dfb_channel_sum: out[c] += sum over pixels of g[:, c] (bias gradient of the un-normalised decoder convolutions).
 *   dfb_upsample2x: backward = 0: in [n,h,w,C] -> out [n,2h,2w,C]; backward = 1: in = grad_out [n,2h,2w,C] -> out [n,h,w,C].
 * f32 = 1 selects fp32 tensors (parity mode) instead of bf16; dfb_channel_sum can also accumulate [2][C] sum / sum of
 * squares in double (stats2) -- the BatchNorm statistics of an fp32 tensor. */
int dfb_bn2d_finalize(const double* stats, double count, int C, int training, float eps, float momentum,
                      const float* gamma, const float* beta, float* running_mean, float* running_var, float* bn,
                      void* stream);
int dfb_bn_gelu_apply(const void* x, const float* bn, int C, long long n_pix, void* y, int f32, void* stream);
int dfb_bn_gelu_backward(const void* x, const void* gy, const float* bn, int C, long long n_pix, int training,
                         double* red, void* gx, float* g_gamma, float* g_beta, float* g_bias, int f32, void* stream);
/* The same in two phases around a caller's all-reduce (SyncBatchNorm, REF/train.py:128): phase 1 = `red` [2][C] from this
 * rank's pixels + the parameter gradients (local sums); phase 2 = gx from `red` summed over all ranks, count_total = pixels
 * per channel over all ranks.  phase 0 = dfb_bn_gelu_backward. */
int dfb_bn_gelu_backward_phase(const void* x, const void* gy, const float* bn, int C, long long n_pix, int training,
                               double* red, void* gx, float* g_gamma, float* g_beta, float* g_bias, int f32, int phase,
                               double count_total, void* stream);
/* Per-channel sums of the data gradient of a 3x3 / stride 1 / pad 1 convolution WITHOUT reading that gradient -- the bias
 * gradient of the producer convolution (REF/src/models/basic/unet.py:25-37): from gy [n,H,W,cout] only its border rows /
 * columns are read, gy_total f32[cout] = sum of gy over all pixels, w f32 torch layout [cout,cin_total,3,3]; colsum f32[cin]
 * for the input channels [cin_off, cin_off + cin).  border_ws: f32[8*cout] scratch. */
int dfb_conv3x3_dgrad_colsum(const void* gy, int f32, int n, int H, int W, int cout, const float* gy_total, const float* w,
                             int cin_total, int cin_off, int cin, float* border_ws, float* colsum, void* stream);
int dfb_channel_sum(const void* g, int C, long long n_pix, float* out, double* stats2, int f32, void* stream);
int dfb_upsample2x(const void* in, int n, int h, int w, int C, void* out, int backward, int f32, void* stream);

/* ------------------------------------------------------------------ per-point decoder stages
 * ConvGRU / ConvGRUDecoder / LinearDecoder (REF/src/models/basic/decoder.py:71-119, 177-253).  The gate and MLP
 * matrices run as 1x1 dfb_conv2d calls over the point list viewed as an [1, n_pad/8, 8, C] image; these are the
 * HBM-bound stages between them.  All bf16 buffers have n_pad rows (n_pad % 8 == 0); rows >= n are written as zero.
 *   dfb_offset_encode:  x[n_pad,cx] bf16 = Linear(3,cx)(offsets[n,3])            (decoder.py:200,228)
 *   dfb_to_bf16_pad:    bf16 zero-padded copy of an fp32 [n,C] matrix
 *   dfb_gru_rh:         rh = sigmoid(r_pre) * h            zr_pre [n_pad,256] bf16 = (z_pre | r_pre), h fp32 [n_pad,128]
 *   dfb_gru_update:     h' = (1 - z) h + z tanh(q_pre)     -> h' fp32 and its bf16 copy
 *   dfb_gru_bwd1/2:     the two elementwise stages of the backward of one GRU iteration (see csrc/gru_elem.cu)
 *   dfb_acc_bf16:       acc fp32 += a (+ b), a / b bf16
 *   dfb_head_out:       flow[n,3] = Linear(32,3)(GELU(y1))  and its backward (dy1, grad W2 [3,32] +=, grad b2 +=)
 *   dfb_offset_encode_backward: grad W [cx,3] += dx^T offsets, grad b += sum dx   (dx fp32 [n_pad,cx])
 * f32 = 1: the gate / head tensors (x, zr_pre, q_pre, rh, dq_pre, dzr_pre, d_rh, y1, dy1, a, b) are fp32 (parity mode). */
int dfb_offset_encode(const float* offs, const float* w, const float* b, int n, int n_pad, int cx, void* x, int f32,
                      void* stream);
int dfb_offset_encode_backward(const float* dx, const float* offs, int n, int cx, float* gw, float* gb, void* stream);
int dfb_to_bf16_pad(const float* src, int n, int n_pad, int C, void* dst, void* stream);
int dfb_gru_rh(const void* zr_pre, const float* h, int n, int n_pad, void* rh, int f32, void* stream);
int dfb_gru_update(const void* zr_pre, const void* q_pre, const float* h, int n, int n_pad, float* h_new, void* hb_new,
                   int f32, void* stream);
int dfb_gru_bwd1(const void* zr_pre, const void* q_pre, const float* h, const float* dh_new, int n, int n_pad,
                 void* dq_pre, void* dzr_pre, float* dh_acc, int f32, void* stream);
int dfb_gru_bwd2(const void* zr_pre, const float* h, const void* d_rh, int n, int n_pad, void* dzr_pre, float* dh_acc,
                 int f32, void* stream);
int dfb_acc_bf16(float* acc, const void* a, const void* b, long long n_elems, int f32, void* stream);
int dfb_head_out(const void* y1, const float* w2, const float* b2, int n, float* flow, int f32, void* stream);
int dfb_head_out_backward(const void* y1, const float* w2, const float* dflow, int n, int n_pad, void* dy1, float* gw2,
                          float* gb2, int f32, void* stream);

/* ------------------------------------------------------------------ fused persistent GRU decoder (tensor cores)
 * ConvGRUDecoder.forward_single for all points at once (REF/src/models/basic/decoder.py:210-237, 184-193): offset
 * encoder, `iters` GRU iterations and the MLP head in ONE kernel; gate weights resident in shared memory, hidden state
 * in registers, gate GEMMs on tcgen05.  wzr bf16 [256][192] (Wz rows, then Wr rows), wq bf16 [128][192], w1 bf16
 * [32][192], K order [h(128), x(64)].  par (fp32): bz[128] br[128] bq[128] b1[32] W2[3*32] b2[3] Woff[64*3] boff[64].
 * h0 bf16 [n_pad,128].  Saved for the backward pass (may be NULL for inference): hsave bf16 [iters+1][n_pad][128]
 * (state entering every iteration, then the final state), xsave bf16 [n_pad][64], y1 bf16 [n_pad][32].
 * dfb_gru_fused_backward: gradient of the `iters` iterations given dh_in / dx_in (bf16, from the head); writes rh, dq
 * [iters][n_pad][128] and dzr [iters][n_pad][256] (bf16 operands of the weight-gradient GEMMs), dh0 bf16 [n_pad][128]
 * and dx fp32 [n_pad][64]; par = bz | br | bq. */
int dfb_gru_fused_forward(const void* h0, const float* offsets, const void* wzr, const void* wq, const void* w1,
                          const float* par, int n, int n_pad, int iters, void* hsave, void* xsave, void* y1,
                          float* flow, void* stream);
int dfb_gru_fused_backward(const void* hsave, const void* xsave, const void* dh_in, const void* dx_in, const void* wzr,
                           const void* wq, const float* par, int n, int n_pad, int iters, void* rh, void* dq, void* dzr,
                           void* dh0, float* dx, void* stream);

/* ------------------------------------------------------------------ input feed
 * Device-side collate_fn_pad (REF/src/dataset.py:22-74; the single-sample strip of ModelWrapper.run_model_wo_ground_data,
 * REF/src/trainer.py:268-282): the raw samples of one frame slot arrive concatenated -- pts f32[total,3], ground u8[total]
 * (1 = ground point, dropped), offs i32[B+1] (device) -- and leave as the padded batch the model consumes:
 * pts_out f32[B,Nmax,3] = pc[~gm] in original order, NaN rows after the kept points; optional per-point payloads that follow
 * the same mask (flow f32[total,3], valid u8, cls u8) -> zero-padded [B,Nmax,...] (pad_sequence default).  keep_counts
 * i32[B] = kept points per sample.  workspace: dfb_collate_workspace(B, max_points_per_sample) ints; its last int is set
 * to 1 if a sample kept more than Nmax points (the surplus is dropped). */
long long dfb_collate_workspace(int B, int max_points_per_sample);
int dfb_collate_pad(const float* pts, const unsigned char* ground, const int* offs, int B, int max_points_per_sample,
                    int Nmax, const float* flow, const unsigned char* valid, const unsigned char* cls, float* pts_out,
                    float* flow_out, unsigned char* valid_out, unsigned char* cls_out, int* keep_counts, int* workspace,
                    void* stream);

/* ------------------------------------------------------------------ evaluation metrics on the device
 * One pass over the points of a validation frame accumulates (+=) everything the reference's three per-frame metric
 * functions reduce on the host: evaluate_leaderboard / evaluate_leaderboard_v2 / evaluate_ssf
 * (REF/src/utils/eval_metric.py:28-106 -> REF/src/utils/av2_eval.py:460-553, 839-870, 872-915).
 * est_flow = final flow (pose flow + estimate), rigid_flow = pose flow, gt_flow: f32[n,3]; pc0: f32 rows of pc_stride floats
 * (xyz first); is_valid, cls: u8[n].  acc: double[DFB_EVAL_ACC_DOUBLES], layout in csrc/eval_metric.cu:
 * [0,8) three-way subset counts, [8,16) their EPE sums, [16,19) TP FP FN, [19,784) bucketed [5 classes][51 speed
 * buckets][count, sum error, sum speed], [784,814) range-wise [5 ranges][static, dynamic][count, sum error, sum distance].
 * tables (host struct, passed by value to the kernel): class id -> {0 background, 1 foreground, 255 neither}
 * (av2_eval.py:217-229), class id -> bucketed meta class row {0..4, 255} (av2_eval.py:47-75, row order of
 * eval_metric.py:262), the speed bucket edges (np.linspace(0, 2, 51) + inf, av2_eval.py:848) and range edges (:892). */
#define DFB_EVAL_ACC_DOUBLES 814
typedef struct dfb_eval_tables {
  unsigned char fg_bg[256];
  unsigned char meta[256];
  double speed_splits[52];
  double dist_splits[6];
  int n_speed, n_dist;
} dfb_eval_tables;
int dfb_eval_accumulate(const float* est_flow, const float* rigid_flow, const float* pc0, int pc_stride,
                        const float* gt_flow, const unsigned char* is_valid, const unsigned char* cls, long long n,
                        const dfb_eval_tables* tables, double* acc, void* stream);

/* ------------------------------------------------------------------ chamfer nearest neighbours (seflowLoss)
 * chamfer3D.forward / .backward (REF/assets/cuda/chamfer3D/chamfer3D_cuda.cpp, kernels chamfer3D.cu:33-124), used by
 * seflowLoss (REF/src/lossfuncs.py:22-100).  pc0 f32[n0,3], pc1 f32[n1,3]; dist0[i] = min_j |pc0[i]-pc1[j]|^2 (squared),
 * idx0[i] = the lowest j attaining it (1e20 / -1 for an empty target); dist1 / idx1 the other way round.
 * workspace: n0 + n1 64-bit words.  backward: grad_pc0 / grad_pc1 f32[n,3] are zeroed inside, then receive both directions. */
int dfb_chamfer_forward(const float* pc0, int n0, const float* pc1, int n1, float* dist0, float* dist1, int* idx0, int* idx1,
                        unsigned long long* workspace, void* stream);
int dfb_chamfer_backward(const float* pc0, int n0, const float* pc1, int n1, const int* idx0, const int* idx1,
                         const float* grad_dist0, const float* grad_dist1, float* grad_pc0, float* grad_pc1, void* stream);

/* ------------------------------------------------------------------ hard voxelisation
 * hard_voxelize_forward(points, voxel_size, coors_range, voxels, coors, num_points_per_voxel, voxel_num, max_points,
 * max_voxels, NDim=3, deterministic) (REF/assets/cuda/mmcv/voxelization.cpp:36-60, voxelization_cuda.cu:8-148): three calls
 *   1. dfb_dynamic_voxelize_forward -> coors i32[n,3];   2. dfb_scatter_index on them -> point2voxel_map, pil_start, sorted_pt;
 *   3. dfb_hard_voxelize_assign: voxels f32[max_voxels,max_points,num_features], voxel_coors i32[max_voxels,3] and
 *      num_points_per_voxel i32[max_voxels] (all pre-zeroed by the caller, as voxelize.py:88-93 does), voxel_num i32[1]
 *      (device).  Voxels are numbered in order of first appearance; a voxel keeps its first max_points points.
 * workspace: dfb_hard_voxelize_workspace(n) ints. */
long long dfb_hard_voxelize_workspace(int n);
int dfb_hard_voxelize_assign(const float* points, int n, int num_features, const int* coors, const int* point2voxel_map,
                             const int* pil_start, const int* sorted_pt, int max_points, int max_voxels, float* voxels,
                             int* voxel_coors, int* num_points_per_voxel, int* voxel_num, int* workspace, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* DEFLOW_B200_H */
