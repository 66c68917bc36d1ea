/*
 * p2r_b200.h -- C ABI of libp2r_b200.so: the B200 (sm_100a) kernels behind the P2RNet hot path.
 *
 * Boundary rules
 *   - plain C: raw DEVICE pointers, int sizes, an opaque `void* stream` (a cudaStream_t / CUstream;
 *     NULL = legacy default stream).  No torch / ATen types.
 *   - every tensor is contiguous, row-major, in the layout the reference operator uses; dtypes are
 *     float32 / int32 / int64 / float64 / uint8 exactly as named below.
 *   - every call is asynchronous on `stream` and performs no allocation and no host sync.
 *   - return value: 0 on success, a positive cudaError_t value for a launch failure, -1 for a bad
 *     argument.  p2r_last_error() returns a thread-local description.  (The reference calls
 *     exit(-1) on a launch failure, _ext-src/include/cuda_utils.h:30-39.)
 *   - outputs marked [zeroed by caller] are accumulated into with atomics, mirroring the
 *     reference's torch::zeros allocation + atomicAdd kernels.
 *
 * Citations "ref:" are into /root/reference.
 */
#ifndef P2R_B200_H
#define P2R_B200_H

#ifdef __cplusplus
extern "C" {
#endif

/* ---- library ------------------------------------------------------------------------------- */
int p2r_abi_version(void);
int p2r_compiled_arch(void); /* 100 */
const char* p2r_last_error(void);
int p2r_device_sm_count(int device, int* sm_count, int* cc_major, int* cc_minor);

/* ---- the nine pointnet2_ops._ext operators --------------------------------------------------
 * ref: external/pointnet2_ops_lib/pointnet2_ops/_ext-src/src/bindings.cpp:6-19 (pybind names),
 *      sampling.cpp:15-87, ball_query.cpp:8-32, group_points.cpp:12-62, interpolate.cpp:14-99.     */

/* furthest_point_sampling(points f32[B,N,3], nsamples) -> i32[B,nsamples]   (sampling.cpp:66-87)
 * scratch: b*n floats, only read when n > 32768 (may be NULL otherwise).                          */
int p2r_furthest_point_sampling(const float* xyz, int b, int n, int m, int* idxs, float* scratch, void* stream);

/* gather_points(points f32[B,C,N], idx i32[B,M]) -> f32[B,C,M]              (sampling.cpp:15-38)  */
int p2r_gather_points(const float* points, const int* idx, int b, int c, int n, int m, float* out, void* stream);

/* gather_points_grad(grad_out f32[B,C,M], idx, N) -> f32[B,C,N] [zeroed by caller] (sampling.cpp:40-64) */
int p2r_gather_points_grad(const float* grad_out, const int* idx, int b, int c, int n, int m, float* grad_points,
                           void* stream);

/* ball_query(new_xyz f32[B,M,3], xyz f32[B,N,3], radius, nsample) -> i32[B,M,nsample]
 * (ball_query.cpp:8-32; every slot is written, no pre-zeroing needed)                             */
int p2r_ball_query(const float* new_xyz, const float* xyz, int b, int n, int m, float radius, int nsample, int* idx,
                   void* stream);

/* group_points(points f32[B,C,N], idx i32[B,P,S]) -> f32[B,C,P,S]           (group_points.cpp:12-36) */
int p2r_group_points(const float* points, const int* idx, int b, int c, int n, int npoints, int nsample, float* out,
                     void* stream);

/* group_points_grad(grad_out f32[B,C,P,S], idx, N) -> f32[B,C,N] [zeroed by caller] (group_points.cpp:38-62) */
int p2r_group_points_grad(const float* grad_out, const int* idx, int b, int c, int n, int npoints, int nsample,
                          float* grad_points, void* stream);

/* three_nn(unknown f32[B,n,3], known f32[B,m,3]) -> (dist2 f32[B,n,3], idx i32[B,n,3])  (interpolate.cpp:14-40)
 * dist2 is the SQUARED distance, like the reference's native op (Python takes the sqrt).         */
int p2r_three_nn(const float* unknown, const float* known, int b, int n, int m, float* dist2, int* idx, void* stream);

/* three_interpolate(points f32[B,C,m], idx i32[B,n,3], weight f32[B,n,3]) -> f32[B,C,n] (interpolate.cpp:42-70) */
int p2r_three_interpolate(const float* points, const int* idx, const float* weight, int b, int c, int m, int n,
                          float* out, void* stream);

/* three_interpolate_grad(grad_out f32[B,C,n], idx, weight, m) -> f32[B,C,m] [zeroed by caller] (interpolate.cpp:71-99) */
int p2r_three_interpolate_grad(const float* grad_out, const int* idx, const float* weight, int b, int c, int n, int m,
                               float* grad_points, void* stream);

/* ---- graph stage (ref: net_utils/vn_dgcnn_util.py) ----------------------------------------- */
/* knn(x f32[B,C,N], k) -> i64[B,N,k]                                         (vn_dgcnn_util.py:4-10) */
int p2r_knn_graph(const float* x, int b, int c, int n, int k, long long* idx, void* stream);
/* get_graph_offset(x f32[B,3d,N], idx i64[B,N,k]) -> f32[B,N,k,d,3] = x[idx]-x  (vn_dgcnn_util.py:70-95) */
int p2r_graph_offset(const float* x, const long long* idx, int b, int d3, int n, int k, float* out, void* stream);

/* 'uniform' seed sampling of STGCN.forward (ref: models/p2rnet/modules/stgcn.py:96-101): hip f32 with `stride`
 * floats between frames (T*stride between sequences) -> seed_inds i64[B,S]                              */
int p2r_uniform_seed_inds(const float* hip, int stride, int b, int t, int s, long long* seed_inds, void* stream);

/* ---- losses (ref: net_utils/nn_distance.py:34-61) ------------------------------------------ */
/* mode 0 = squared L2, 1 = L1 (l1=True), 2 = smooth-L1 (l1smooth=True, delta).                    */
int p2r_nn_distance(const float* pc1, const float* pc2, int b, int n, int m, int c, int mode, float delta,
                    float* dist1, long long* idx1, float* dist2, long long* idx2, void* stream);
/* backward of the above; g1 f32[B,N] / g2 f32[B,M] may be NULL; grads [zeroed by caller].        */
int p2r_nn_distance_grad(const float* pc1, const float* pc2, const long long* idx1, const long long* idx2,
                         const float* g1, const float* g2, int b, int n, int m, int c, int mode, float delta,
                         float* grad_pc1, float* grad_pc2, void* stream);

/* ---- eval post-processing (ref: net_utils/ap_helper.py:133-255, nms.py, box_util.py) -------- */
/* decode: center f32[B,K,3], log_size f32[B,K,3], heading (sin,cos) f64[B,K,2], hip f32 with
 * `hip_stride` floats between frames and T*hip_stride between scenes ->
 * corners f64[B,K,8,3] (utils/tools.py:33-51 order), aabb f64[B,K,6] (min xyz, max xyz),
 * nonempty u8[B,K] (size in [0.01,10] and some hip point inside the box enlarged by `contact`).  */
int p2r_decode_boxes(const float* center, const float* log_size, const double* heading_sincos, const float* hip,
                     int hip_stride, int b, int k, int t, double contact, double* corners, double* aabb,
                     unsigned char* nonempty, void* stream);
/* nms_3d_faster / nms_3d_faster_samecls (nms.py:41-119): boxes f64[B,K,6], score f64[B,K],
 * valid u8[B,K] or NULL, cls i32[B,K] or NULL -> keep u8[B,K], order i32[B,K] (-1 padded).        */
int p2r_nms3d(const double* boxes, const double* score, const unsigned char* valid, const int* cls, int b, int k,
              double thr, int old_type, unsigned char* keep, int* order, void* stream);
/* box3d_iou (box_util.py:90-118) for every pair: c1 f64[P,8,3], c2 f64[G,8,3] -> f64[P,G] x2.    */
int p2r_box3d_iou(const double* corners1, const double* corners2, int np_, int ng, double* iou3d, double* iou2d,
                  void* stream);

/* ---- dense per-point layers, channel-last (rows = points, columns = channels) -----------------
 * dtype codes: 0 = float32, 1 = bfloat16 (storage; arithmetic is fp32, BN column sums are fp64).
 * Replace the cuDNN/cuBLAS calls behind the reference's 1x1 Conv / BatchNorm / ReLU stacks:
 * ref: models/p2rnet/modules/sub_modules.py:88-113 (SingleConv 'cbr'), stgcn.py:45-67,
 *      stgcn_layers.py:50-67,402-414, vote_center.py:28-32, proposal_net.py:63-94,
 *      pointnet2_modules.py:9-19,243-247.                                                          */

/* C[M,N] (+)= op(A).op(B) (+bias[N]) (ReLU).  trans_a: A stored [K,M]; trans_b: B stored [N,K]
 * (the nn.Linear / 1x1-conv weight layout).  splits > 1: split-K, fp32 atomics into zeroed C.      */
int p2r_sgemm(int M, int N, int K, const void* A, int lda, int trans_a, int a_dtype, const void* B, int ldb,
              int trans_b, int b_dtype, void* C, int ldc, int c_dtype, const float* bias, int relu, int accumulate,
              int splits, void* stream);
/* column sums of x[M,C]: s1 = sum x, s2 = sum x^2 (f64[C], zeroed by caller)                       */
int p2r_col_stats(const void* x, int dtype, long long M, int C, double* s1, double* s2, void* stream);
/* backward column sums: dz = dy masked by the ReLU (relu = 0 none, 1: y > 0, 2: x*scale+shift > 0, recomputed);
 * s1 = sum dz, s2 = sum dz*(x-mean)*rstd (s2/x may be NULL)                                                       */
/* relu: 0 none, 1 mask from y > 0, 2 mask recomputed from x*scale+shift > 0, 3 `y` points to the bit mask written by
 * p2r_affine_act (streaming kernels only) -- same meaning in p2r_bn_bwd_apply                                      */
int p2r_col_bwd_stats(const void* dy, const void* x, const void* y, int dtype, long long M, int C, const float* mean,
                      const float* rstd, int relu, double* s1, double* s2, const float* scale, const float* shift,
                      void* stream);
/* column sums of a wide [M,C] matrix (any C % VEC == 0), e.g. the 1600-wide graph-conv bias gradient; s1 zeroed by caller */
int p2r_col_sum_wide(const void* dy, int dtype, long long M, int C, double* s1, void* stream);

/* training-mode BatchNorm statistics -> mean, rstd, fused scale/shift; updates running stats like torch.
 * s1 / s2 may be `copies` partial sums, `copy_stride` doubles apart (the GEMM-epilogue statistics of p2r_gemm_bf16_ex) */
int p2r_bn_finalize(int C, long long M, const double* s1, const double* s2, int copies, long long copy_stride,
                    const float* gamma, const float* beta, float eps, float momentum, float* running_mean,
                    float* running_var, float* mean, float* rstd, float* scale, float* shift, void* stream);
/* y = x*scale[c] + shift[c] (+residual) (ReLU).  relu_mask (optional, only where p2r_stream_bn_supported(...) == 1):
 * [M, C/8] bytes, bit i of byte j = (y[row, 8 j + i] > 0) -- the backward passes read it (relu mode 3 below) instead of
 * the whole output tensor.                                                                          */
int p2r_affine_act(const void* x, int dtype, long long M, int C, const float* scale, const float* shift,
                   const void* residual, int relu, void* y, unsigned char* relu_mask, void* stream);
/* 1 when [M, C] operands of this dtype take the bulk-TMA streaming kernels (bf16, C = 64, M >= 4096)              */
int p2r_stream_bn_supported(int dtype, long long M, int C);
/* BN(+ReLU)(+residual) backward, elementwise part (s1 == NULL: eval-mode BN).  colsum / period (optional, only where
 * p2r_stream_bn_supported(...) == 1, period <= 32): also accumulate colsum[row % period][c] += dx[row][c] (double,
 * zero-filled by the caller) -- the bias gradient of the graph convolution whose output rows cycle through the joints
 * (stgcn_layers.py:50-56 bias), without another pass over dx.                                       */
int p2r_bn_bwd_apply(const void* dy, const void* x, const void* y, int dtype, long long M, int C, const float* mean,
                     const float* rstd, const float* scale, const double* s1, const double* s2, int relu, void* dx,
                     void* dres, const float* shift, double* colsum, int period, void* stream);
/* The same launch; sums64 / sums32 (both or neither): the finished [2][C] double sums of p2r_col_bwd_stats (d beta,
 * d gamma) are also stored as float32 [2][C] -- the parameters' gradient dtype without a conversion launch.          */
int p2r_bn_bwd_apply_ex(const void* dy, const void* x, const void* y, int dtype, long long M, int C, const float* mean,
                        const float* rstd, const float* scale, const double* s1, const double* s2, int relu, void* dx,
                        void* dres, const float* shift, double* colsum, int period, const double* sums64, float* sums32,
                        void* stream);
/* dz = dy * (y > 0)                                                                                */
int p2r_relu_bwd(const void* dy, const void* y, int dtype, long long total, void* dz, void* stream);
/* (KT x 1) temporal conv as GEMM: x[B,T,V,C] -> col[B*T*V, KT*C] (zero padded), and its adjoint    */
int p2r_temporal_unfold(const void* x, int dtype, int B, int Tn, int V, int C, int KT, void* col, void* stream);
int p2r_temporal_fold(const void* dcol, int dtype, int B, int Tn, int V, int C, int KT, void* dx, void* stream);
/* channel-last grouping + max-pool of the SA layer (ref: pointnet2_utils.py:319-346, pointnet2_modules.py:243-247) */
int p2r_group_rows(const void* feats, int dtype, const int* idx, int B, int N, int C, int P, int S, void* out,
                   void* stream);
int p2r_group_rows_grad(const void* grad, int dtype, const int* idx, int B, int N, int C, int P, int S, float* dfeats,
                        void* stream);
/* Adjoint of out[b][p][:] = feats[b][idx[b][p]][:] (p2r_group_rows with S = 1: the seed-frame pick before conv_joint,
 * ref: models/p2rnet/modules/stgcn.py:136-139) written destination-major into dfeats[B,N,C] of the gradient's dtype:
 * every row is stored (zeros where no slot picked it, the sum where several did), no zero-fill, no atomics.          */
int p2r_select_rows_grad(const void* grad, int dtype, const int* idx, int B, int N, int C, int P, void* dfeats,
                         void* stream);
int p2r_maxpool_rows(const void* x, int dtype, long long R, int S, int C, void* out, unsigned char* arg, void* stream);
int p2r_maxpool_rows_grad(const void* dout, int dtype, const unsigned char* arg, long long R, int S, int C, void* dx,
                          void* stream);
/* The set-abstraction layer's group -> shared MLP (two 1x1 convs + ReLU) -> max over nsample as ONE tcgen05 kernel
 * (ref: PointnetSAModuleVotes.forward, pointnet2_modules.py:220-256, with QueryAndGroup's grouping_operation,
 * pointnet2_utils.py:319-346; csrc/sa_fused.cu): the (B, C, P, S) grouped tensor and the MLP activations are never
 * written to memory.  feats bf16 [B,N,C] channel-last, idx i32 [B,P,S], w1 / w2 bf16 [C,C] (Conv2d 1x1 weights), b1 / b2
 * f32 [C] or NULL; C = 256; S a power of two <= 128.  out [B*P, C] (out_dtype 0 = f32, 1 = bf16) = max_s relu(w2 .
 * relu(w1 . feats[idx[.., s]] + b1) + b2); argmax u8 [B*P, C] or NULL (first maximum, for the backward pass); h1 bf16
 * [B*P*S, C] or NULL: when given, the first activation is also stored (training keeps it for the backward pass).   */
int p2r_sa_fused(const void* feats, const int* idx, const void* w1, const float* b1, const void* w2, const float* b2,
                 int b, int n, int p, int s, int c, void* out, int out_dtype, unsigned char* argmax, void* h1,
                 void* stream);

/* x[f,j,:] = sk[f,j,:] + mean_k pos[f,k,:] and its backward dpos[f,k,:] = (1/K) sum_j dx[f,j,:]
 * (ref: models/p2rnet/modules/stgcn.py:121,129)                                                                 */
int p2r_embed_sum(const void* sk, const void* pos, int dtype, long long frames, int J, int K, int C, void* x,
                  void* stream);
int p2r_embed_sum_grad(const void* dx, int dtype, long long frames, int J, int K, int C, void* dpos, void* stream);

/* K <= 4 linear layers (3 -> 64 first layers, ref: stgcn.py:45-50) as memory-bound maps: y = x.W^T (+bias) and
 * dW[N,K] (f32, zeroed by the caller) = dz^T.x.  N % VEC == 0, (256*VEC) % N == 0 (VEC = 4 f32 / 8 bf16).          */
int p2r_smallk_linear(const void* x, const float* W, const float* bias, int dtype, long long M, int N, int K, void* y,
                      void* stream);
int p2r_smallk_dw(const void* dz, const void* x, int dtype, long long M, int N, int K, float* dW, void* stream);
/* The same two maps with float32 x (point coordinates are never rounded to bf16) and bf16 y / dz: throughput mode. */
int p2r_smallk_linear_mixed(const float* x, const float* W, const float* bias, long long M, int N, int K, void* y,
                            void* stream);
int p2r_smallk_dw_mixed(const void* dz, const float* x, long long M, int N, int K, float* dW, void* stream);

/* First layer of a point MLP (Conv1d K<=4 -> N, BatchNorm, ReLU; ref: stgcn.py:45-50, sub_modules.py:88-113) without its
 * pre-activation ever being stored (csrc/embed_ops.cu): z = W x is linear in the coordinates, so the BatchNorm statistics
 * follow from the moments of x, and forward / backward recompute z from 12 bytes instead of moving [M, N] tensors.
 *   p2r_coord_moments      s f64[K + K*K] (zero-filled) += (sum x_k, sum x_i x_j)
 *   p2r_embed_l1_finalize  moments -> mean / rstd / scale / shift f32[N] (+ running statistics, like p2r_bn_finalize)
 *   p2r_embed_l1_fwd       y bf16[M,N] = relu(scale (W x) + shift)
 *   p2r_embed_l1_bwd_stats s1 f64[N] += sum g, s2 += sum g xhat   (g = dy masked by the ReLU; dy bf16[M,N])
 *   p2r_embed_l1_bwd_dw    dW f32[N,K] (zero-filled) += sum_m dz_m x_m^T, dz = scale (g - s1/M - xhat s2/M) (s1 = NULL: eval)
 * x f32[M,K], W f32[N,K]; N a multiple of 8 that divides 2048.                                                          */
int p2r_coord_moments(const float* x, long long M, int K, double* s, void* stream);
int p2r_embed_l1_finalize(const double* s, long long M, int K, const float* W, int N, const float* gamma, const float* beta,
                          float eps, float momentum, float* running_mean, float* running_var, float* mean, float* rstd,
                          float* scale, float* shift, void* stream);
int p2r_embed_l1_fwd(const float* x, const float* W, const float* scale, const float* shift, long long M, int N, int K,
                     void* y, void* stream);
int p2r_embed_l1_bwd_stats(const void* dy, const float* x, const float* W, const float* mean, const float* rstd,
                           const float* scale, const float* shift, long long M, int N, int K, double* s1, double* s2,
                           void* stream);
int p2r_embed_l1_bwd_dw(const void* dy, const float* x, const float* W, const float* mean, const float* rstd,
                        const float* scale, const float* shift, const double* s1, const double* s2, long long M, int N,
                        int K, float* dW, void* stream);

/* bf16 tensor-core GEMM (tcgen05 + TMA + TMEM), the throughput-mode backend of every dense layer, above all the
 * fused graph-convolution GEMM that replaces conv 64->704 + einsum (ref: stgcn_layers.py:58-67).
 * C[M,N] (+)= op(A).op(B)^T, fp32 accumulate.  a_mn = 0: A is [M,K] row-major, 1: [K,M]; b_mn = 0: B is [N,K]
 * row-major (nn.Linear weight layout), 1: [K,N].  c_dtype 0 = fp32, 1 = bf16.  splits > 1: split-K with fp32
 * atomics into zeroed C.  block_n in {0 = auto, 64, 128, 160, 256}.  Pointers 16-byte aligned, pitches % 8 == 0. */
int p2r_gemm_bf16(int M, int N, int K, const void* A, int lda, int a_mn, const void* B, int ldb, int b_mn, void* C,
                  int ldc, int c_dtype, const float* bias, int relu, int splits, int block_n, void* stream);
/* The same GEMM with the three extras the graph convolution of st_gcn_block uses (stgcn_layers.py:58-67 computes
 * conv 64 -> 11*64 then einsum 'nkctv,kvw->nctw'; here it is ONE GEMM against W_eff = sum_k W_k (x) A_k):
 *   kb_list [tiles_n][kb_stride] int: entry 0 = how many 64-wide k-blocks n-tile i visits, entries 1.. their indices --
 *     W_eff is zero in every 64x64 block whose joints are further apart than the adjacency's max hop;
 *   tile_mask [tiles_m][tiles_n] bytes: 0 = structurally-zero output tile, skipped (C left untouched; weight gradient);
 *   stats [stat_copies][2][64] double, zeroed by the caller: per channel (column % 64) sum and sum of squares of the
 *     stored output = the statistics of the BatchNorm that follows (stgcn_layers.py:403), so it needs no extra pass.
 * Extras need splits <= 1 and an explicit block_n (the tables are indexed by this launch's tiling); NULL = off.     */
int p2r_gemm_bf16_ex(int M, int N, int K, const void* A, int lda, int a_mn, const void* B, int ldb, int b_mn, void* C,
                     int ldc, int c_dtype, const float* bias, int relu, int splits, int block_n, const int* kb_list,
                     int kb_stride, const unsigned char* tile_mask, double* stats, int stat_copies, void* stream);

/* CTA-pair (tcgen05 cta_group::2, cluster of two CTAs on one TPC) variant for the large K-major GEMMs of the graph
 * convolution: C[M,N] bf16 = A[M,K] . B[N,K]^T (+bias)(ReLU), one 256 x block_n tile per pair (block_n 128 / 256),
 * each CTA staging only half of the B tile.  kb_list / stats as in p2r_gemm_bf16_ex, n-tiles of width block_n.       */
int p2r_gemm_bf16_pair(int M, int N, int K, const void* A, int lda, const void* B, int ldb, void* C, int ldc,
                       const float* bias, int relu, int block_n, const int* kb_list, int kb_stride, double* stats,
                       int stat_copies, void* stream);

/* Weight gradient of the graph convolution on CTA pairs: dW[N1,N2] fp32 (zero-filled by the caller) += dz^T . x with
 * dz [R,N1], x [R,N2] bf16 row-major (reduction over the rows).  tile_list = (row tile, column tile) int pairs of the
 * 256 x 256 output tiles to compute (the structurally non-zero ones); the reduction is split `splits` ways and the
 * partial tiles are combined by bulk-tensor reduce-add stores.                                                      */
int p2r_gemm_bf16_pair_dw(int R, int N1, int N2, const void* dz, int ldz, const void* x, int ldx, float* dW, int ldw,
                          const int* tile_list, int num_tiles, int splits, void* stream);

/* (KT x 1) temporal convolution (zero padding (KT-1)/2) as an implicit tensor-core GEMM over the 3-D activation tensor
 * [B, rows = T*V, C], a tap shifting by V rows; no unfold buffer (ref: st_gcn_block.tcn conv, stgcn_layers.py:405-411).
 * mode 0: y = conv(x, W2[Co, KT*Ci]) (+bias); mode 1: dx from dy and Wt[KT*Co, Ci]; mode 2: dW2[Co, KT*Ci] (fp32,
 * zeroed by the caller when splits > 1) from x (act) and dy (other).  rows % 128 == 0, Ci = Co = 64.
 * stats (mode 0 only, optional): [stat_copies][2][64] double, zeroed by the caller -- per-channel sum / sum of squares
 * of y, i.e. the statistics of the BatchNorm2d that follows the conv (stgcn_layers.py:412).                        */
int p2r_tconv_bf16(int mode, const void* act, const void* w, const void* other, void* out, int B, int rows, int Ci,
                   int Co, int KT, int V, const float* bias, int splits, double* stats, int stat_copies, void* stream);
/* AdamW over n float32 tensors in a few launches, capturable in a CUDA graph: the update rule of torch.optim.AdamW
 * (amsgrad = False, maximize = False) that the reference's optimiser factory builds (models/optimizers.py:90).
 * params / grads / exp_avg / exp_avg_sq: HOST arrays of n device pointers, numel: host array of n element counts
 * (positive: float32 tensors; negative: -count elements of float64 -- parameter, gradient and moments alike);
 * step: DEVICE float = number of updates already applied (read by every launch, incremented once at the end).    */
int p2r_adamw_step(int n, const void* const* params, const void* const* grads, void* const* exp_avg,
                   void* const* exp_avg_sq, const long long* numel, float* step, double lr, double beta1, double beta2,
                   double eps, double weight_decay, void* stream);
/* Diagnostic (not part of the reference's interface): per-tile globaltimer stamps of one CTA of the halo temporal-conv
 * kernels into device_buffer[3 roles][64 tiles][8 events] (long long, device memory); NULL switches it off again. */
int p2r_debug_tconv_trace(long long* device_buffer);

/* Weight plumbing of the fused graph convolution (ref: ConvTemporalGraphical.forward, stgcn_layers.py:58-67; A = adjacency
 * stack * edge importance, stgcn.py:133-134).  Build  W_eff[(w,co),(v,ci)] = sum_k A[k,v,w] W[k*Co+co, ci]  (bf16), its
 * transpose (bf16, the operand of the input-gradient GEMM) and  b_eff[(w,co)] = sum_k b[k*Co+co] sum_v A[k,v,w]  (fp32)
 * from conv_w [K*Co, Ci] fp32, conv_b [K*Co] fp32 (or NULL), A [K,V,V] fp32.  Co = Ci = 64.                          */
int p2r_gcn_build_weight(const float* conv_w, const float* conv_b, const float* A, int K, int V, int Co, int Ci,
                         void* w_eff, void* w_eff_t, float* b_eff, void* stream);
/* ... and its backward: fold dW_eff [V*Co, V*Ci] fp32 and db_eff [V*Co] fp32 onto d_conv_w [K*Co, Ci], d_conv_b [K*Co]
 * and dA [K,V,V] (all three written in full, no atomics: deterministic; dA is 0 where A == 0).                    */
int p2r_gcn_reduce_weight_grad(const float* dw_eff, const float* db_eff, const float* conv_w, const float* conv_b,
                               const float* A, int K, int V, int Co, int Ci, float* d_conv_w, float* d_conv_b,
                               float* dA, void* stream);

/* ---- sample -> batch (ref: models/p2rnet/dataloader.py) ------------------------------------------------------------
 * One launch builds the three large tensors of the `data` dict for a whole batch from raw samples RESIDENT in device
 * memory: frame picking (dataloader.py:128-131), flip / rotate / translate of joints and votes (augment_data,
 * dataloader.py:31-84, bit-exact incl. the reference's float32 / float64 rounding points -- csrc/augment_math.h), dtype
 * casts (:137-145) and batching (collate_fn, :148-160).
 *   joints f32 [F_total, J, 3], votes f32 [F_total, J, 10] (column 0 = vote mask): all raw frames of all samples, packed;
 *   frame_start i64 [n_samples + 1]: first packed frame of each sample; sample_ids i32 [B]: the batch;
 *   params f64 [B, 16]: per batch item (enabled, flip, rot[9] row-major, shift[3], floor height, 0) -- the host draws
 *   them with the reference's RNG calls (dataloader.py:33-37); enabled = 0 copies the raw values (val / test);
 *   out_channels 3, or 4 to append joint height above the floor (use_height, dataloader.py:112-115).
 * Outputs: input_joints f32 [B,T,J,out_channels], vote_label f32 [B,T,J,9], vote_label_mask i64 [B,T,J]; T = num_frames. */
int p2r_make_batch(const float* joints, const float* votes, const long long* frame_start, const int* sample_ids,
                   const double* params, int b, int num_frames, int j, int out_channels, float* input_joints,
                   float* vote_label, long long* vote_label_mask, void* stream);
/* Diagnostic twin of p2r_make_batch with the data-movement variant chosen by the caller: 1 = one CTA per 8 output frames
 * (P2R_MAKE_BATCH_VARIANT=1 makes p2r_make_batch use it), 2 = persistent CTAs with a 3-stage cp.async ring (the default
 * since round 2: 2.2x faster on a B200).  Same arithmetic, bit-identical outputs; bench.py times one against the other.                              */
int p2r_make_batch_variant(int variant, const float* joints, const float* votes, const long long* frame_start,
                           const int* sample_ids, const double* params, int b, int num_frames, int j, int out_channels,
                           float* input_joints, float* vote_label, long long* vote_label_mask, void* stream);

/* ---- detection loss (ref: models/loss.py:42-189 BoxNetDetectionLoss) ------------------------------------------------
 * The whole loss in one launch (csrc/loss_ops.cu, arithmetic in csrc/loss_math.h): vote loss (compute_vote_loss, :90-115),
 * proposal <-> ground-truth correspondence + objectness (compute_correspondence, :117-150), centre / size / heading /
 * class losses (compute_box_and_sem_cls_loss, :42-88) and the statistics of __call__ (:152-189).
 * Predictions: vote_xyz f32 [B,S,3]; center, size, agg (aggregated_vote_xyz) f32 [B,P,3]; heading [B,P,2] float64 when
 * heading_f64 else float32; obj f32 [B,P,2] and sem f32 [B,P,C] with ROW strides obj_stride / sem_stride in elements (the
 * two are slices of one [B,P,2+C] tensor in the model); skeleton (seed_skeleton) f32 [B,S,J,3]; seed_inds i64 [B,S].
 * Ground truth (the `data` dict, dataloader.py:137-146): vote_label f32 [B,T,J,9], vote_mask i64 [B,T,J], gt_center /
 * gt_size f32 [B,G,3], gt_mask f32 [B,G], gt_heading f32 [B,G,2], gt_cls i64 [B,G]; origin = origin_joint_id.
 * Outputs: out32 f32 [8] = vote, objectness, center, size, sem_cls loss, pos_ratio, neg_ratio, obj_acc; out64 f64 [2] =
 * heading loss, total (float64 in the reference too: the heading mixture is float64); scales f64 [4] and the
 * un-normalised gradients u_* (shapes of the matching predictions; u_c1 / u_c2: the two halves of the centre loss) for
 * p2r_detection_loss_grad.  workspace: p2r_detection_loss_workspace(b, s) doubles, ZEROED by the caller.            */
long long p2r_detection_loss_workspace(int b, int s);
int p2r_detection_loss(const float* vote_xyz, const float* center, const float* size, const void* heading,
                       int heading_f64, const float* obj, int obj_stride, const float* sem, int sem_stride,
                       const float* agg, const float* skeleton, const long long* seed_inds, const float* vote_label,
                       const long long* vote_mask, const float* gt_center, const float* gt_mask, const float* gt_size,
                       const float* gt_heading, const long long* gt_cls, int b, int s, int j, int t, int p, int g, int c,
                       int origin, float* out32, double* out64, double* scales, float* u_vote, float* u_c1, float* u_c2,
                       float* u_size, void* u_head, float* u_obj, float* u_sem, double* workspace,
                       long long workspace_doubles, void* stream);
/* backward: g32 f32 [8] / g64 f64 [2] = upstream gradients of out32 / out64 -> gradients of vote_xyz, center, size,
 * heading (dtype as in the forward), obj [B,P,2] and sem [B,P,C] (both contiguous), every element written.            */
int p2r_detection_loss_grad(const float* g32, const double* g64, const double* scales, const float* u_vote,
                            const float* u_c1, const float* u_c2, const float* u_size, const void* u_head,
                            int heading_f64, const float* u_obj, const float* u_sem, int b, int s, int p, int c,
                            float* d_vote, float* d_center, float* d_size, void* d_head, float* d_obj, float* d_sem,
                            void* stream);

/* ---- Gaussian-mixture box heads (ref: models/p2rnet/modules/mdn.py:31-84 MixtureDensityHead) ---------------------------
 * Training-time point prediction with n_samples = 1 (the configuration proposal_net.py:141-147 builds):
 *   out[r,:] = sum_g sigmoid(logits[r,g]) * (mu[g,:] + exp(log_sigma[g,:]) * eps[r,g,:])
 * logits [rows,G] float32, or bfloat16 when logits_bf16 (the output of the pi 1x1 conv); mu [G,d] and eps [rows,G,1,d]
 * float64 when mu_f64 (the heading head, whose mu grid is float64 in the reference) else float32; log_sigma f32 [G,d];
 * out [rows,d] in mu's type.  eps ~ N(0,1) is drawn by the caller with the reference's torch call (mdn.py:44).
 * G <= 256, d <= 4.  csrc/gmm_ops.cu, arithmetic in csrc/gmm_math.h.                                                 */
int p2r_gmm_mix(const void* logits, int logits_bf16, const void* mu, int mu_f64, const float* log_sigma, const void* eps,
                long long rows, int g, int d, void* out, void* stream);
/* backward: dout [rows,d] (mu's type) -> dlogits [rows,G] (logits' type), dmu [G,d] (mu's type), dls f32 [G,d], all
 * written in full (deterministic two-level sum over rows).  workspace: p2r_gmm_mix_workspace(rows, g, d) doubles,
 * ZEROED by the caller.                                                                                             */
long long p2r_gmm_mix_workspace(long long rows, int g, int d);
int p2r_gmm_mix_grad(const void* logits, int logits_bf16, const void* mu, int mu_f64, const float* log_sigma,
                     const void* eps, const void* dout, long long rows, int g, int d, void* dlogits, void* dmu, float* dls,
                     double* workspace, long long workspace_doubles, void* stream);

/* ---- vote tail (ref: models/p2rnet/modules/vote_center.py:52-58 + network.py:89-90) -------------------------------------
 * vote_xyz = seed_xyz + net[:, 0:3];  vote_feat = v / ||v||_2 with v = seed_feat + net[:, 3:]  in one launch.
 * net [rows, 3+C] float32, or bfloat16 when net_bf16 (the output of the last voting conv); seed_xyz f32 with xyz_stride
 * floats between rows (the hip joint of seed_skeleton, read in place); seed_feat f32 [rows, C].
 * Outputs: vote_xyz f32 [rows,3], vote_feat f32 [rows,C] (L2-normalised), norm f32 [rows] (kept for the backward).      */
int p2r_vote_tail(const void* net, int net_bf16, const float* seed_xyz, long long xyz_stride, const float* seed_feat,
                  long long rows, int c, float* vote_xyz, float* vote_feat, float* norm, void* stream);
/* backward: g_xyz f32 [rows,3] / g_feat f32 [rows,C] (either may be NULL = zero) -> d_net [rows,3+C] in net's type and
 * d_seed_feat f32 [rows,C], every element written.                                                                  */
int p2r_vote_tail_grad(const float* g_xyz, const float* g_feat, const float* vote_feat, const float* norm,
                       long long rows, int c, void* d_net, int net_bf16, float* d_seed_feat, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* P2R_B200_H */
