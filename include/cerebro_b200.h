/*
 * cerebro_b200.h -- C ABI of the B200-native loop-detection hot path.
 *
 * Drop-in boundary for the three seams of mpkuse/cerebro's loop-closure path
 * (reference file:line cited per entry point):
 *
 *   descriptor  <-  ROS service /whole_image_descriptor_compute
 *                   (srv/WholeImageDescriptorCompute.srv:1-5, server
 *                   scripts/whole_image_desc_compute_server.py:596-650, client
 *                   src/Cerebro.cpp:243-275)
 *   index       <-  faiss::IndexFlatIP add/search/ntotal (src/Cerebro.cpp:390-460) and the
 *                   three GEMVs + arg-max of Cerebro::descrip_N__dot__descrip_0_N
 *                   (src/Cerebro.cpp:1019-1043)
 *   pnp         <-  StaticTheiaPoseCompute::PNP (src/DlsPnpWithRansac.h:169-179,
 *                   src/DlsPnpWithRansac.cpp:132-245)
 *
 * Conventions
 *   - every function returns 0 on success, a negative CB_E* code on failure; no C++
 *     exception, abort or exit crosses this boundary.  cb_last_error() returns a
 *     thread-local message for the last failure on the calling thread.
 *   - handles are opaque; each handle owns its device memory and one CUDA stream and is
 *     meant to be driven by ONE host thread (the reference drives each seam from its own
 *     thread: desc_th, dot_product_th, loopcandidate_consumer_th, cerebro_node.cpp:487-509).
 *     Different handles may be used concurrently from different threads.
 *   - "_device" variants take raw device pointers (memory owned by the caller, e.g. a
 *     torch tensor's data_ptr) and a cudaStream_t passed as void* (NULL = the CUDA default
 *     stream, as everywhere in CUDA); they enqueue on exactly that stream and do not
 *     synchronise.  Use one stream per handle for its device calls (scratch is per handle).  All other variants take
 *     HOST pointers, copy in and out on the handle's own stream, and return when the result is in the caller's buffer.
 *     An index handle orders the two against each other: whenever consecutive calls launch on different streams, the
 *     later stream first waits (cudaStreamWaitEvent) for the work the earlier call enqueued, so a host search after an
 *     add_device / search_device on the caller's stream sees the rows, planes and scratch that call produced.  The
 *     caller's stream must still be alive at the next call on the handle.
 *   - there is no CPU fallback anywhere behind this ABI: without a CUDA device every
 *     create call fails with CB_ENODEVICE.
 */
#ifndef CEREBRO_B200_H
#define CEREBRO_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(__GNUC__)
#define CB_API __attribute__((visibility("default")))
#else
#define CB_API
#endif

#define CB_OK 0
#define CB_EINVAL (-1)    /* bad argument */
#define CB_ENODEVICE (-2) /* no usable CUDA device / wrong architecture */
#define CB_ECUDA (-3)     /* CUDA runtime error, see cb_last_error() */
#define CB_ENOMEM (-4)    /* capacity exceeded / allocation failed */
#define CB_EREFUSED (-5)  /* input refused by the algorithm (reference returns -1) */

CB_API int cb_version(void);             /* ABI version, currently 2 (1 -> 2: vlad_ghost, cb_comm_*, cb_index_search_sharded*) */
CB_API const char* cb_last_error(void);  /* thread-local, never NULL */
CB_API int cb_device_count(void);        /* number of visible CUDA devices (0 if none) */

/* ------------------------------------------------------------------------------------
 * index: the keyframe descriptor database (row-major fp32, one row per keyframe)
 * ---------------------------------------------------------------------------------- */
typedef struct cb_index cb_index;

/* tie rule when two rows have exactly the same score */
#define CB_TIE_LOW_LABEL 0  /* lower label wins (stable order, FAISS-like)            */
#define CB_TIE_HIGH_LABEL 1 /* higher label wins: src/Cerebro.cpp:1039-1043 keeps the   */
                            /* LAST index equal to the max                             */

/* IndexFlatIP(d) (src/Cerebro.cpp:390).  `capacity` rows are preallocated on `device`
 * (the reference preallocates 29000 columns, src/Cerebro.cpp:946).  A sharded index
 * holds the rows whose global label g satisfies g % world == rank; world=1, rank=0 is the
 * plain single-GPU index.  Labels are always GLOBAL insertion order. */
CB_API int cb_index_create(cb_index** out, int d, int64_t capacity, int device, int rank, int world);
CB_API int cb_index_destroy(cb_index* ix);
CB_API int cb_index_reset(cb_index* ix);
CB_API int64_t cb_index_ntotal(const cb_index* ix); /* GLOBAL number of rows added (all shards) */
CB_API int64_t cb_index_nlocal(const cb_index* ix); /* rows resident on this shard */
CB_API int cb_index_dim(const cb_index* ix);

/* index.add(n, x) (src/Cerebro.cpp:431).  x holds n consecutive GLOBAL rows [n][d]; a
 * sharded index keeps only its own.  _f64 takes the VectorXd the reference stores in
 * DataNode (src/DataNode.cpp:427-444) and narrows to fp32 on the device. */
CB_API int cb_index_add(cb_index* ix, int64_t n, const float* x);
CB_API int cb_index_add_f64(cb_index* ix, int64_t n, const double* x);
CB_API int cb_index_add_device(cb_index* ix, int64_t n, const float* x_dev, void* stream);

/* Bulk-load rows straight into THIS shard's local storage (DataManager::loadStateFromDisk,
 * src/DataManager.cpp:1304-1320, re-lists stored descriptors on resume): local row r gets the
 * global label r * world + rank.  Do not mix with cb_index_add on a sharded index. */
CB_API int cb_index_add_local_device(cb_index* ix, int64_t n_local, const float* x_dev, void* stream);

/* Optional device-side timing of the HBM sweep kernel (CUDA events on the launching stream
 * around every scores_kernel launch, up to 64 launches between reads).  get returns the summed
 * duration and the number of launches since the last read, and resets both. */
CB_API int cb_index_set_timing(cb_index* ix, int on);
CB_API int cb_index_get_sweep_timing(cb_index* ix, double* total_ms, int* n_launches);

/* index.search(nq, xq, k, distances, labels) (src/Cerebro.cpp:460): the k largest inner
 * products per query in descending order, labels = global insertion index, padded with
 * (-inf, -1).  Only rows with label < limit_rows take part (limit_rows < 0: all rows):
 * this is the reference's 50/150-descriptor lag (src/Cerebro.cpp:914,1019 / :374,415).
 * 1 <= k <= 32.  The sweep accumulates in fp32; the 32 best rows per query are then
 * re-scored in fp64 on the device, so ordering and `scores_f64` match an fp64 reference.
 * `scores_f64` may be NULL. */
CB_API int cb_index_search(cb_index* ix, int nq, const float* xq, int k, int64_t limit_rows, int tie_mode,
                    float* distances, int64_t* labels, double* scores_f64);

/* Same on device buffers; outputs [nq][k] fp64 scores and int64 labels of THIS shard. */
CB_API int cb_index_search_device(cb_index* ix, int nq, const float* xq_dev, int k, int64_t limit_rows,
                           int tie_mode, double* scores_dev, int64_t* labels_dev, void* stream);

/* Merge `n_lists` top-k lists laid out [n_lists][nq][k_in] (e.g. the all-gathered
 * per-shard results) into [nq][k_out], same ordering rule.  Pure device function. */
CB_API int cb_topk_merge_device(int n_lists, int nq, int k_in, const double* scores_dev,
                         const int64_t* labels_dev, int k_out, int tie_mode, double* out_scores_dev,
                         int64_t* out_labels_dev, void* stream);

/* ---- sharded database over the GPUs of one box (BASELINE configs 3 and 4).  The reference has no
 * collective anywhere (its only transport is ROS); this is the scaling axis the B200 build adds: DB rows round-robin over
 * `world` ranks (one per GPU), every rank scans its shard for all queries, ONE ncclAllGather of the per-shard top-k lists,
 * and the same deterministic merge on every rank.  A rank is a process (torchrun) or a thread of one process (a ROS node
 * driving 8 GPUs); NCCL is loaded at run time (dlopen), a single-GPU host never touches it. */
typedef struct cb_comm cb_comm;
#define CB_COMM_ID_BYTES 128 /* sizeof(ncclUniqueId) */
/* rank 0 makes the id and hands it to the other ranks by whatever means the host has (a ROS parameter, a file,
 * torch.distributed ...); then every rank calls cb_comm_create with the same id (collective: returns when all have). */
CB_API int cb_comm_get_unique_id(uint8_t* id_out /* [CB_COMM_ID_BYTES] */);
CB_API int cb_comm_create(cb_comm** out, const uint8_t* id, int rank, int world, int device);
CB_API int cb_comm_destroy(cb_comm* c);
CB_API int cb_comm_rank(const cb_comm* c);
CB_API int cb_comm_world(const cb_comm* c);
CB_API int cb_comm_nccl_version(void); /* e.g. 22809, -1 when NCCL cannot be loaded */

/* Attach a communicator to a sharded index (rank / world must equal the index's). */
CB_API int cb_index_attach_comm(cb_index* ix, cb_comm* c);

/* Collective search: every rank passes ITS OWN nq_local queries (the descriptors of the keyframes it just computed; the
 * same nq_local on every rank).  One call = ncclAllGather of the queries -> local sweep + top-k of all world * nq_local
 * queries over this shard -> ONE ncclAllGather of the packed per-shard lists (fp64 score, int64 label) -> merge.
 * Outputs the merged GLOBAL top-k of this rank's own queries: scores [nq_local][k] fp64, labels [nq_local][k] int64.
 * Deterministic and independent of the sharding: the lists carry fp64 re-scored values (see cb_index_search).
 * cb_index_gathered_queries returns the device block [world * nq_local][d] the call gathered (rank-major: global order of
 * the new keyframes) -- passing it to cb_index_add_device appends the step's descriptors to the sharded DB without a
 * second exchange. */
CB_API int cb_index_search_sharded_device(cb_index* ix, int nq_local, const float* xq_local_dev, int k, int64_t limit_rows,
                                   int tie_mode, double* scores_dev, int64_t* labels_dev, void* stream);
/* host buffers in and out (blocking), the call a C++ / ROS host makes */
CB_API int cb_index_search_sharded(cb_index* ix, int nq_local, const float* xq_local, int k, int64_t limit_rows, int tie_mode,
                            float* distances, int64_t* labels, double* scores_f64);
CB_API const float* cb_index_gathered_queries(const cb_index* ix);

/* One iteration of Cerebro::descrip_N__dot__descrip_0_N (src/Cerebro.cpp:1019-1081) for
 * list length l on a NON-sharded index that already holds rows [0,l): scores of rows
 * l-1, l-2, l-3 against rows [0, l-lag), arg-max with the last-index tie rule, locality
 * and threshold test.  out[0]=1 if a loop candidate was found else 0; out_prev = arg-max
 * of the newest descriptor; out_score = its fp64 score; argmax3 = the three arg-maxes. */
CB_API int cb_index_naive_candidate(cb_index* ix, int64_t l, int lag, int locality_thresh,
                             float dot_thresh, int* out_found, int64_t* out_prev,
                             double* out_score, int64_t argmax3[3]);

/* Copy rows [first, first+n) of this shard's LOCAL storage to the host (testing and
 * DataManager::saveStateToDisk, src/DataManager.cpp:1157-1168). */
CB_API int cb_index_get_rows(cb_index* ix, int64_t first_local, int64_t n, float* out);

/* raw device pointer to local row storage (row-major, stride d floats) */
CB_API const float* cb_index_device_rows(const cb_index* ix);

/* ------------------------------------------------------------------------------------
 * descriptor: NetVLAD whole-image descriptor forward pass
 * ---------------------------------------------------------------------------------- */
typedef struct cb_descriptor cb_descriptor;

/* Network description handed over by the host side (weights already BN-folded, fp32):
 * MobileNet-v1 prefix + NetVLAD as listed in scripts/keras.models/model.json. */
typedef struct cb_netvlad_weights {
  int in_channels;        /* 1 or 3 */
  int n_blocks;           /* depthwise/pointwise blocks after conv1 */
  const float* conv1_w;   /* [3][3][in_channels][32]  (kh,kw,cin,cout) */
  const float* conv1_b;   /* [32] */
  const float* const* dw_w; /* n_blocks x [3][3][C]  */
  const float* const* dw_b; /* n_blocks x [C]        */
  const float* const* pw_w; /* n_blocks x [C][Cout]  (NULL entry: block has no pointwise) */
  const float* const* pw_b; /* n_blocks x [Cout]     */
  const int* dw_stride;   /* n_blocks, 1 or 2 */
  const int* channels_out; /* n_blocks, Cout of the block */
  int vlad_k;             /* clusters the soft-assignment runs over (NetVLADLayer: K; GhostVLADLayer: K + ghosts); <= 16 */
  int vlad_d;             /* feature dim entering NetVLAD */
  const float* vlad_w;    /* [D][K] */
  const float* vlad_b;    /* [K] */
  const float* vlad_c;    /* [D][K] */
  int vlad_ghost;         /* GhostVLADLayer (scripts/predict_utils.py:110-141): the LAST vlad_ghost clusters take part in the
                             softmax but are dropped before the two norms (:133); descriptor dim = (vlad_k - vlad_ghost) * D.
                             0 = NetVLADLayer */
} cb_netvlad_weights;

/* HDF5ModelImageDescriptor.__init__(kerasmodel_file, im_rows, im_cols, im_chnls)
 * (server.py:492-593): allocates activations for `max_batch` frames of rows x cols x chnls. */
CB_API int cb_descriptor_create(cb_descriptor** out, const cb_netvlad_weights* w, int rows, int cols,
                         int chnls, int max_batch, int device);
CB_API int cb_descriptor_destroy(cb_descriptor* d);
/* The June2019 models (scripts/keras.models/June2019/ ...mobilenetv2-block_9_add..., selected at
 * launch/mynteye_vinsfusion.launch:100 and built by keras_helpers.py from keras_applications'
 * MobileNetV2): Conv1 3x3 s2 (pad bottom/right) + ReLU6, then inverted-residual blocks
 * [expand 1x1 + ReLU6] -> depthwise 3x3 (s1 'same' | pad bottom/right + s2 'valid') + ReLU6 ->
 * project 1x1 (linear) [+ block input], then NetVLAD.  Weights BN-folded, fp32, natural
 * (unpadded) channel counts, multiples of 8. */
typedef struct cb_ir_block {
  int c_in, c_exp, c_out;  /* block input / expanded / output channels */
  int stride;              /* depthwise stride, 1 or 2 */
  int residual;            /* 1: output = project + block input (the Keras Add layer) */
  const float* expand_w;   /* [c_in][c_exp], NULL for expanded_conv (then c_exp == c_in) */
  const float* expand_b;   /* [c_exp] */
  const float* dw_w;       /* [3][3][c_exp] */
  const float* dw_b;       /* [c_exp] */
  const float* project_w;  /* [c_exp][c_out] */
  const float* project_b;  /* [c_out] */
} cb_ir_block;

typedef struct cb_netvlad_v2_weights {
  int in_channels;       /* 1 or 3 */
  const float* conv1_w;  /* [3][3][in_channels][32] */
  const float* conv1_b;  /* [32] */
  int n_blocks;
  const cb_ir_block* blocks;
  int vlad_k, vlad_d;
  const float* vlad_w; /* [D][K] */
  const float* vlad_b; /* [K] */
  const float* vlad_c; /* [D][K] */
  int vlad_ghost;      /* as in cb_netvlad_weights */
} cb_netvlad_v2_weights;

/* Same handle type and the same compute / dim / destroy calls as cb_descriptor_create. */
CB_API int cb_descriptor_create_v2(cb_descriptor** out, const cb_netvlad_v2_weights* w, int rows, int cols,
                            int chnls, int max_batch, int device);

CB_API int cb_descriptor_dim(const cb_descriptor* d); /* K * D, what the probe call learns (Cerebro.cpp:113-120) */

/* handle_req (server.py:596-650) for `n` frames: uint8 images [n][rows][cols][chnls]
 * (row_stride_bytes between image rows, 0 = tight) -> fp32 descriptors [n][K*D], unit L2 norm,
 * index k*D+d.  The ROS shim widens to float64[] (srv:4). */
CB_API int cb_descriptor_compute(cb_descriptor* d, int n, const uint8_t* images, int64_t row_stride_bytes,
                          float* out);
/* The same, writing the service reply's `float64[] desc` (srv/WholeImageDescriptorCompute.srv:4, server.py:648; the client
 * copies it into a VectorXd, Cerebro.cpp:268-271) directly: fp32 crosses the bus, the widening happens into `out`.
 * Both calls accept pageable `images` (a cv::Mat's data: Cerebro.cpp:243-256) as well as pinned memory: pageable frames go
 * through a pinned staging ring inside the library (CPU copy of the next chunk overlaps the DMA + forward pass of the
 * current one), pinned frames are uploaded in place. */
CB_API int cb_descriptor_compute_f64(cb_descriptor* d, int n, const uint8_t* images, int64_t row_stride_bytes,
                              double* out);
CB_API int cb_descriptor_compute_device(cb_descriptor* d, int n, const uint8_t* images_dev, float* out_dev,
                                 void* stream);
/* debugging / parity: copy the activation after layer `layer` (0 = conv1, then dw1, pw1, dw2, ...)
 * of frame 0 as fp32 NHWC to the host. Returns number of floats written or negative error. */
CB_API int64_t cb_descriptor_get_activation(cb_descriptor* d, int layer, float* out, int64_t max_floats);

/* ------------------------------------------------------------------------------------
 * pnp: batched DLS-PnP RANSAC
 * ---------------------------------------------------------------------------------- */
typedef struct cb_pnp cb_pnp;

typedef struct cb_ransac_params { /* theia::RansacParameters as set at DlsPnpWithRansac.cpp:207-212 */
  double error_thresh;        /* 0.03 */
  double min_inlier_ratio;    /* 0.7  */
  int max_iterations;         /* 50 (hypotheses evaluated per candidate) */
  int min_iterations;         /* 5  */
  int use_mle;                /* 1  */
  double failure_probability; /* 0.01 (Theia default) */
  int adaptive;               /* 1: Theia's sequential adaptive termination is replayed over the
                                    hypotheses; 0: all max_iterations hypotheses count (BASELINE config 5) */
  uint64_t seed;              /* counter-based sampler key (used when samples == NULL) */
} cb_ransac_params;

CB_API void cb_ransac_params_default(cb_ransac_params* p);

CB_API int cb_pnp_create(cb_pnp** out, int max_candidates, int max_points_total, int max_hypotheses, int device);
CB_API int cb_pnp_destroy(cb_pnp* p);

/* StaticTheiaPoseCompute::PNP for a batch of loop candidates.
 *   offsets [n_cand+1]: candidate c owns correspondences [offsets[c], offsets[c+1])
 *   X  [total][3] fp64 : 3-D points in frame a            (w_X,  DlsPnpWithRansac.cpp:132)
 *   uv [total][2] fp64 : normalised image coords in b     (c_uv_normalized)
 *   samples: NULL, or int32 [n_cand][max_iterations][15] indices local to the candidate
 *   c_T_w [n_cand][16] row-major 4x4, confidence [n_cand] (= return value; -1 when a
 *   candidate has < 20 points, :136-139), num_iterations, n_inliers, best_hyp (may be NULL). */
CB_API int cb_pnp_solve_batch(cb_pnp* p, int n_cand, const int32_t* offsets, const double* X,
                       const double* uv, const cb_ransac_params* params, const int32_t* samples,
                       double* c_T_w, float* confidence, int32_t* num_iterations,
                       int32_t* n_inliers, int32_t* best_hyp);
CB_API int cb_pnp_solve_batch_device(cb_pnp* p, int n_cand, const int32_t* offsets_dev, int total_points,
                              const double* X_dev, const double* uv_dev, const cb_ransac_params* params,
                              const int32_t* samples_dev, double* c_T_w_dev, float* confidence_dev,
                              int32_t* num_iterations_dev, int32_t* n_inliers_dev,
                              int32_t* best_hyp_dev, void* stream);

/* StaticTheiaPoseCompute::P3P_ICP (src/DlsPnpWithRansac.cpp:16-122): RANSAC over
 * AlignPointCloudsUmeyamaWithRansac (src/DlsPnpWithRansac.h:117-166; SampleSize 10, model accepted iff
 * min(s,1/s) > 0.9, Error = ||R a + t - b||).  A, B [total][3] fp64: the same points in frames a and b.
 * The caller sets params->error_thresh = 0.1 (DlsPnpWithRansac.cpp:89).  samples: NULL or
 * int32 [n_cand][max_iterations][10].  Outputs as cb_pnp_solve_batch (b_T_a row-major 4x4). */
CB_API int cb_pnp_icp_batch(cb_pnp* p, int n_cand, const int32_t* offsets, const double* A, const double* B,
                     const cb_ransac_params* params, const int32_t* samples, double* b_T_a,
                     float* confidence, int32_t* num_iterations, int32_t* n_inliers, int32_t* best_hyp);
CB_API int cb_pnp_icp_batch_device(cb_pnp* p, int n_cand, const int32_t* offsets_dev, const double* A_dev,
                            const double* B_dev, const cb_ransac_params* params, const int32_t* samples_dev,
                            double* b_T_a_dev, float* confidence_dev, int32_t* num_iterations_dev,
                            int32_t* n_inliers_dev, int32_t* best_hyp_dev, void* stream);

/* Minimal solver only (theia::DlsPnp as called at DlsPnpWithRansac.h:61) for n_sets sets of
 * exactly `m` points: up to 27 solutions per set.  n_solutions [n_sets]; R [n_sets][27][9]
 * row-major, t [n_sets][27][3].  For parity tests of the solver. */
CB_API int cb_pnp_dls_minimal(cb_pnp* p, int n_sets, int m, const double* X, const double* uv,
                       int32_t* n_solutions, double* R, double* t);
/* parity tests: intermediates of the last cb_pnp_dls_minimal call, per set -- what = 0: the 27 x 27 action matrix
 * (729 doubles), 1: the 60 gradient coefficients.  Returns the number of doubles copied, negative on error. */
CB_API int64_t cb_pnp_debug_read(cb_pnp* p, int what, int n_sets, double* out, int64_t max_doubles);

/* ------------------------------------------------------------------------------------
 * frontend: ORB descriptors + keypoints + depth images of a loop candidate's two frames ->
 * Hamming matches -> GMS inliers -> the 3D-2D / 3D-3D sets the pose solvers take
 * (StaticPointFeatureMatching, src/utils/PointFeatureMatching.cpp; ORB detection and the
 * stereo block matcher stay OpenCV's on the host)
 * ---------------------------------------------------------------------------------- */
typedef struct cb_frontend cb_frontend;

CB_API int cb_frontend_create(cb_frontend** out, int max_pairs, int max_features /* per image, <= 16384 */, int device);
CB_API int cb_frontend_destroy(cb_frontend* f);

/* cv::BFMatcher(cv::NORM_HAMMING).match(d1, d2) (PointFeatureMatching.cpp:40-43) followed by
 * gms_matcher(kp1, size1, kp2, size2, matches).GetInlierMask(mask, false, false) (:50-52;
 * src/utils/GMSMatcher/gms_matcher.cpp:5-181) for a batch of image pairs.
 *   off1, off2 [n_pairs+1]: pair p owns query features [off1[p], off1[p+1]) and train features
 *                           [off2[p], off2[p+1]) of the concatenated arrays
 *   desc1, desc2 : [total][32] bytes, 256-bit ORB descriptors; kp1, kp2: [total][2] float KeyPoint.pt (x, y) in pixels
 * One match per query descriptor, in query order (what match() returns):
 *   train_idx [total1] index local to the pair's train set (first minimum wins; -1 if the pair has no train features),
 *   distance [total1] Hamming distance (may be NULL), inlier_mask [total1] 0/1, n_inliers [n_pairs]. */
CB_API int cb_frontend_match_gms(cb_frontend* f, int n_pairs, const int32_t* off1, const int32_t* off2,
                          const uint8_t* desc1, const uint8_t* desc2, const float* kp1, const float* kp2,
                          int width1, int height1, int width2, int height2, int32_t* train_idx,
                          int32_t* distance, uint8_t* inlier_mask, int32_t* n_inliers);
/* device time (ms) of the matching + GMS kernels of the last cb_frontend_match_gms call (CUDA events) */
CB_API float cb_frontend_last_match_ms(const cb_frontend* f);

/* The correspondence sets of the batch matched last on this handle, from its GMS inliers, in match order:
 *   mode 0: make_3d_2d_collection__using__pfmatches_and_disparity (PointFeatureMatching.cpp:96-153):
 *           X = 3-D point of frame a looked up at ((int)v, (int)u) in img3d_a, kept iff 0.1 <= z <= 25;
 *           uv / uv_d = K^-1 [u v 1]^T of the feature in frame a / b (normalised image coordinates)
 *   mode 1: make_3d_3d_collection__using__pfmatches_and_disparity (:158-196): X from img3d_a, Y from img3d_b, both gated
 *   mode 2: mode 0 with the frames' roles swapped -- the Option-B call of src/Cerebro.cpp:1562-1565: X = 3-D point of
 *           frame b looked up at ((int)v_d, (int)u_d) in img3d_b (img3d_a may be NULL); uv / uv_d keep their meaning
 *           (frame a / frame b), so Option B solves PNP(X, uv) for a_T_b
 *   img3d_a, img3d_b: [n_pairs][rows][cols][3] float32 (the CV_32FC3 "3d image" of the stereo pair); K_inverse 3x3 row-major.
 * Outputs are laid out per pair at the pair's query offset: pair p's kept entries are rows
 * [off1[p], off1[p] + counts[p]) of X [total1][3], uv / uv_d [total1][2], Y [total1][3]. */
CB_API int cb_frontend_collect(cb_frontend* f, int mode, const float* img3d_a, const float* img3d_b, int rows, int cols,
                        const double* K_inverse, int32_t* counts, double* X, double* uv, double* uv_d, double* Y);

/* Stereo depth of the two frames (src/utils/CameraGeometry.cpp): `bm = cv::StereoBM::create(ndisp, wsz)` (:81, the
 * reference uses 64, 21) and `bm->compute(left, right, disparity)` (:410-418) with OpenCV's default parameters, for n
 * rectified pairs: left, right [n][rows][cols] uint8 -> disparity [n][rows][cols] int16 = disparity * 16, -16 where the
 * matcher rejects the pixel or outside the valid ROI.  Bit-exact against cv2.StereoBM (tests/test_stereo.py). */
CB_API int cb_frontend_stereo_bm(cb_frontend* f, int n, const uint8_t* left, const uint8_t* right, int rows, int cols,
                          int ndisp, int wsz, int16_t* disparity);
CB_API float cb_frontend_last_stereo_ms(const cb_frontend* f); /* device time of the last cb_frontend_stereo_bm */
/* StereoGeometry::disparity_to_3DPoints (CameraGeometry.cpp:459-520): the CV_32FC3 "3d image"
 * ((j + Q03) pw, (i + Q13) pw, Q23 pw), pw = 1 / (disparity / 16 * Q32 + Q33 + 1e-6); Q?? = entries of the 4x4 Q matrix. */
CB_API int cb_frontend_disparity_to_3d(cb_frontend* f, int n, const int16_t* disparity, int rows, int cols, float Q03,
                                float Q13, float Q23, float Q32, float Q33, float* out3d);

/* ------------------------------------------------------------------------------------
 * features: the image side of a loop candidate -- undistortion / stereo-rectification warps and ORB extraction
 * (StereoGeometry::do_image_undistortion / do_stereo_rectification_of_undistorted_images, src/utils/CameraGeometry.cpp:42,
 * 381-382; StaticPointFeatureMatching::gms_point_feature_matches, src/utils/PointFeatureMatching.cpp:16-22)
 * ---------------------------------------------------------------------------------- */
typedef struct cb_features cb_features;

/* One handle per camera geometry: images are rows x cols, 8-bit, one channel; up to max_images per call and max_keypoints
 * keypoints per image (cv::ORB keeps ties at the response threshold, so allow a margin over n_features). */
CB_API int cb_features_create(cb_features** out, int rows, int cols, int max_images, int max_keypoints, int device);
CB_API int cb_features_destroy(cb_features* f);

/* The CV_32FC1 map pair of one warp (camodocal's initUndistortRectifyMap, CameraGeometry.cpp:20-34, or
 * cv::initUndistortRectifyMap, :348-349 -- computed once on the host at start-up), kept on the device in `slot` 0..3. */
CB_API int cb_features_set_remap(cb_features* f, int slot, const float* map_x, const float* map_y);
/* cv::remap(src, dst, map_x, map_y, CV_INTER_LINEAR) (border constant 0) for n images [n][rows][cols]; with slot_b >= 0 the
 * result is warped again through slot_b (raw -> undistorted -> rectified, 8-bit rounding in between exactly like the reference's two
 * calls).  Bit-exact against cv::remap (tests/test_orb.py). */
CB_API int cb_features_remap(cb_features* f, int n, const uint8_t* src, int slot_a, int slot_b, uint8_t* dst);

/* cv::ORB::create(n_features) -> setFastThreshold(0) -> detectAndCompute(image, noArray(), keypoints, descriptors) for n images
 * [n][rows][cols].  Per image i: n_keypoints[i] keypoints in cv::ORB's output order at [i][0 .. n_keypoints[i]):
 *   kp_xy [n][max_keypoints][2] KeyPoint.pt, kp_size, kp_angle (degrees), kp_response (Harris), kp_octave, descriptors
 *   [n][max_keypoints][32].  kp_size / kp_angle / kp_response / kp_octave may be NULL.
 * Keypoints (order, coordinates, size, angle, response, octave) are bit-exact against the installed OpenCV; descriptors agree
 * except for isolated bits (~1 in 10^6) where OpenCV's own 7x7 Gaussian -- an IPP float filter whose rounding depends on the
 * host CPU's code path -- lands on the other side of a .5 tie. */
CB_API int cb_features_orb(cb_features* f, int n, const uint8_t* images, int n_features, int32_t* n_keypoints, float* kp_xy,
                    float* kp_size, float* kp_angle, float* kp_response, int32_t* kp_octave, uint8_t* descriptors);
CB_API float cb_features_last_orb_ms(const cb_features* f); /* stream time of the last cb_features_orb, host selections included */
/* parity tests: pyramid (what = 0), FAST score map (1) or blurred pyramid (2) of image 0 of the last cb_features_orb call, the 8
 * levels back to back; returns the number of bytes written or a negative error */
CB_API int64_t cb_features_debug_read(cb_features* f, int what, uint8_t* out, int64_t max_bytes);

#ifdef __cplusplus
}
#endif
#endif /* CEREBRO_B200_H */
