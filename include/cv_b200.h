/* circuitvision_b200 — C ABI of the B200-native hot path (libcv_b200.so).
 *
 * The reference (JKc66/CircuitVision) is pure Python and has no FFI: its boundary for this path is the pair
 * of Python call signatures
 *     src/sam2_infer.py:191-275        SAM2ImageWrapper.forward (+ SAM2Transforms :29-128, MultiKernelRefinement :130-189)
 *     src/circuit_analyzer.py:321-386  CircuitAnalyzer.segment_with_sam2
 *     src/circuit_analyzer.py:1286-1605 CircuitAnalyzer.get_node_connections
 * which circuitvision_b200/sam2_infer.py and circuitvision_b200/circuit_analyzer.py keep unchanged.  This header
 * is the C ABI *underneath* those classes: plain C symbols, int status returns (0 = ok, message via
 * cv_last_error()), caller-owned DEVICE pointers unless a parameter says "host", explicit CUDA stream passed as
 * void* (cudaStream_t), no torch / C++ types.  INTEGRATION.md shows the ctypes binding.
 */
#ifndef CV_B200_H
#define CV_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define CV_OK 0
#define CV_ERR_INVALID 1
#define CV_ERR_CUDA 2
#define CV_ERR_CAPACITY 3

/* thread-local description of the last non-zero status returned on this thread */
const char* cv_last_error(void);
/* library / build identification: "circuitvision_b200 <version> sm_100a" */
const char* cv_version(void);
/* 1 when the current device is compute capability 10.x */
int cv_device_is_sm100(int device);

/* ------------------------------------------------------------------ node / connection analysis
 * replaces circuit_analyzer.py:1286-1605 (get_node_connections) and helpers :787-809, :461-477, :289-311,
 * :388-412, :811-846 for a batch of B masks of identical size H x W.                                      */

#define CV_BOX_ZERO_IN_MASK 1 /* class not in ('crossover','junction','circuit','vss')  (:1326) */
#define CV_BOX_IS_COMPONENT 2 /* class not in non_components                           (:51,:1381) */
#define CV_BOX_IS_SOURCE 4    /* class in source_components                            (:52,:1408) */
#define CV_BOX_IS_TERMINAL 8  /* class == 'terminal'                                   (:2273)     */

typedef struct cv_box {
  int32_t xmin, ymin, xmax, ymax;     /* int()-truncated mask-space coords            (:1339-1340) */
  int32_t rxmin, rymin, rxmax, rymax; /* resized-space coords int(v*scale)            (:466-469)  */
  int32_t flags;                      /* CV_BOX_*                                                  */
  int32_t thresh;                     /* contact threshold 20 / 8 / 6                 (:1404-1415) */
  int32_t uid_group;                  /* index (within the image) of the first box with the same persistent_uid */
  int32_t reserved;
} cv_box;

typedef struct cv_contour {
  int32_t start_x, start_y;           /* raster-first pixel = first vertex                          */
  int32_t offset, nverts;             /* slice of the image's point pool                            */
  int32_t xmin, ymin, xmax, ymax;     /* inclusive extents; cv2.boundingRect = (xmin,ymin,xmax-xmin+1,ymax-ymin+1) */
  int64_t a00, a01;                   /* polygon sums: contourArea = |a00|/2, m01/m00 = a01/(3*a00) */
  int32_t new_id;                     /* id in new_nodes_list, or -1 when dropped     (:1547-1582) */
  int32_t ncomp;                      /* attached components after uid de-duplication               */
  int32_t has_source;                 /* any attached component is a source                         */
  int32_t centroid_y;                 /* int(m01/m00), INT32_MIN when m00 == 0                      */
} cv_contour;

typedef struct cv_pair {              /* one accepted (node, component) attachment, in reference order */
  int32_t contour, box, px, py;
} cv_pair;

#define CV_STATUS_EXTERNAL_OVERFLOW 1
#define CV_STATUS_CONTOUR_OVERFLOW 2
#define CV_STATUS_POINT_OVERFLOW 4
#define CV_STATUS_PAIR_OVERFLOW 8
#define CV_STATUS_BOX_HITS_OVERFLOW 16 /* reserved: not raised any more (boxes beyond the 64 staged contacts take a recompute path) */

typedef struct cv_image_result {
  int32_t n_external;  /* external components before the area filter */
  int32_t n_contours;  /* kept contours (ids 0..n-1, cv2 order)      */
  int32_t n_points;    /* vertices stored in the point pool          */
  int32_t n_pairs;
  int32_t n_nodes;     /* len(new_nodes_list)                        */
  int32_t ground;      /* old contour id chosen as node 0, or -1     */
  int32_t inverted;    /* get_contours took the mean>127 branch      */
  int32_t status;      /* CV_STATUS_* bits; non-zero => results for this image are incomplete */
} cv_image_result;

typedef struct cv_nodes_caps {
  int32_t max_external; /* candidates per image before the area filter (default 32768) */
  int32_t max_contours; /* kept contours per image              (default 2560)  */
  int32_t max_points;   /* vertices per image                   (default 262144) */
  int32_t max_pairs;    /* attachments per image                (default 8192)  */
} cv_nodes_caps;

/* resized width the reference uses: int(600 * (W / H)) */
int cv_nodes_resized_width(int H, int W);
/* bytes of device scratch cv_nodes_analyze needs for this problem size */
size_t cv_nodes_workspace_bytes(int B, int H, int W, const cv_nodes_caps* caps);

/* All pointers are device pointers.  boxes: concatenated per image, box_offsets[B+1].
 * outputs: emptied [B,H,W] u8 ; resized [B,600,w'] u8 (pre-enhance, used for final_node_viz) ;
 *          enhanced [B,600,w'] u8 (exactly the array the reference returns, incl. its 255->1 mutation) ;
 *          contours [B,max_contours] ; points [B,max_points,2] i32 ; pairs [B,max_pairs] ; results [B].      */
int cv_nodes_analyze(const uint8_t* masks, int B, int H, int W, const cv_box* boxes, const int32_t* box_offsets,
                     int max_boxes_per_image, uint8_t* emptied, uint8_t* resized, uint8_t* enhanced,
                     cv_contour* contours, int32_t* points, cv_pair* pairs, cv_image_result* results,
                     const cv_nodes_caps* caps, void* workspace, size_t workspace_bytes, void* stream);

/* Compacts the used prefixes of the fixed-capacity result tables of cv_nodes_analyze into one contiguous blob, so that
 * only bytes that carry information cross PCIe (bench.py e2e; circuitvision_b200/pipeline.py).  All device pointers.
 * header: int64 [4*(B+1)] = {byte offset of image b's section, n_contours, n_pairs, n_points}, header[4*B] = total bytes;
 * section b = cv_contour[n_contours] | cv_pair[n_pairs] | int32 (x,y)[n_points], 16-byte aligned.  An image whose
 * section would end beyond blob_bytes is not copied (the caller compares header[4*B] with the capacity).              */
int cv_nodes_pack(const cv_contour* contours, const int32_t* points, const cv_pair* pairs, const cv_image_result* results,
                  int B, const cv_nodes_caps* caps, long long* header, uint8_t* blob, long long blob_bytes, void* stream);

/* ------------------------------------------------------------------ terminal reclassification
 * replaces the numeric part of circuit_analyzer.py:2217-2310 (reclassify_terminals_based_on_connectivity) with its
 * helpers segment_circuit :313-319, get_contours(area_threshold=0.0001) :388-412 and is_point_near_bbox(.., 10) :811-846,
 * called from analysis_pipeline.py:127 on the full RGB page: B pages [B,H,W,3] u8 ->
 *   wire_mask [B,H,W] u8   the adaptive-threshold mask after box masking (prelim_wire_mask, :2238-2249)
 *   box_counts [n_boxes_total] i32   distinct contours with a vertex near the box (-1 for boxes without CV_BOX_IS_TERMINAL);
 *                                    the caller relabels terminals with a count >= 2 as 'voltage.dc' (:2291)
 *   contours / points / results      the page's external contours at native resolution (same tables as cv_nodes_analyze;
 *                                    only n_external, n_contours, n_points, inverted and status are meaningful).
 * cv_box uses xmin..ymax (page coordinates) and flags CV_BOX_ZERO_IN_MASK | CV_BOX_IS_TERMINAL; H <= 12000.           */
size_t cv_terminals_workspace_bytes(int B, int H, int W, const cv_nodes_caps* caps);
int cv_terminals_analyze(const uint8_t* pages_rgb, int B, int H, int W, const cv_box* boxes, const int32_t* box_offsets,
                         int max_boxes_per_image, int n_boxes_total, uint8_t* wire_mask, int32_t* box_counts,
                         cv_contour* contours, int32_t* points, cv_image_result* results, const cv_nodes_caps* caps,
                         void* workspace, size_t workspace_bytes, void* stream);

/* Native-resolution connected-component labelling (BASELINE.json cfg 4; SURVEY §8(d): not a reference code
 * path — oracle is cv2.connectedComponents up to renaming).  labels[p] = 1 + min linear index of p's component,
 * 0 for background.  connectivity 4 or 8.  n_components[B] optional (may be NULL).                        */
size_t cv_ccl_workspace_bytes(int B, int H, int W);
int cv_ccl_label(const uint8_t* masks, int B, int H, int W, int connectivity, int32_t* labels,
                 int32_t* n_components, void* workspace, size_t workspace_bytes, void* stream);
/* Number of kernels the last cv_* call on this thread launched (for bench.py's gpu_launches). */
int cv_last_launch_count(void);

/* ------------------------------------------------------------------ tensor-core building blocks
 * bf16 GEMM on tcgen05 + TMA with fp32 accumulation in TMEM:  C[M,N] = act(A[M,K] * W[N,K]^T + bias) + residual.
 * Replaces every nn.Linear / 1x1 nn.Conv2d on the SAM 2.1 path (sam2_infer.py:226-232,252 via the sam2 package).
 * A, W: bf16 device pointers with row pitch lda/ldw (elements, multiples of 8); N % 32 == 0; K % 8 == 0.
 * act: 0 none, 1 exact-erf GELU, 2 ReLU.  out_f32 and/or out_bf16 (either may be NULL, not both).            */
int cv_gemm_bf16(const void* A, long long lda, const void* W, long long ldw, int M, int N, int K, const float* bias,
                 int act, const float* residual, long long ld_res, float* out_f32, long long ld_f32, void* out_bf16,
                 long long ld_bf16, void* stream);

/* The same GEMM with every epilogue the SAM 2.1 engine uses (the entry the parity tests drive; cv_sam2_forward calls the same
 * launcher).  map_mode: 0 identity, 1 window un-partition (Hiera window_unpartition), 2 ConvTranspose2d(k=2,s=2) pixel shuffle,
 * 3 2x2 max-pool of window-major rows into the pooled grid (Q-pool shortcut, do_pool(proj(x))), 4 qkv of a Q-pooled block: the
 * first pool_cols columns are 2x2 max-pooled into pool_out, the rest stored as 16-bit rows.  operand_fp16: A / W / 16-bit
 * outputs are IEEE half instead of bf16.  res_row_mod > 0: residual row = destination row % res_row_mod.                    */
typedef struct cv_gemm_epilogue {
  const float* bias;
  int act;            /* 0 none, 1 GELU, 2 ReLU */
  int res_before_act; /* 1: act(acc + bias + residual) */
  const float* residual;
  long long ld_res;
  long long res_row_mod;
  float* out_f32;
  long long ld_f32;
  void* out_16;
  long long ld_16;
  int map_mode;
  int ws, nwx, nwy, H, W, cout;
  int pool_cols;
  void* pool_out;
  long long ld_pool;
  int operand_fp16;
} cv_gemm_epilogue;
int cv_gemm_ex(const void* A, long long lda, const void* W, long long ldw, int M, int N, int K, const cv_gemm_epilogue* epilogue,
               void* stream);

/* Fused Hiera MLP half-block (sam2 MultiScaleBlock: x + mlp(norm2(x)), called from sam2_infer.py:226 through the sam2
 * package):  X[M,C] (fp32, in place)  <-  X + fc2(GELU(fc1(LayerNorm(X; gamma, beta, eps)))).  W1 [4C,C] and W2 [C,4C] are
 * 16-bit K-major weights in the operand format selected by operand_fp16, b1 [4C] / b2 [C] fp32.  The normalised operand,
 * the hidden activation and the fc2 accumulator stay on the SM (shared memory / TMEM).  C in {96,112,144,192,224,288}.   */
int cv_mlp_fused(float* X, int M, int C, const float* gamma, const float* beta, float eps, const void* W1, const float* b1,
                 const void* W2, const float* b2, int operand_fp16, void* stream);

/* Debug: a zeroed device buffer of >= 4001 uint64 that the following cv_mlp_fused calls fill with the pipeline timeline of
 * CTA 0 (record = (event * 4096 + index) << 44 | globaltimer ns; buffer[0] = number of records); NULL switches it off. */
int cv_mlp_fused_set_trace(void* device_buffer);
/* same for the global-attention kernel (scripts/attn_trace.py): 4001 uint64, NULL switches it off */
int cv_attn_set_trace(void* device_buffer);
/* same for k_gemm_tc (scripts/gemm_trace.py): a zeroed buffer of 1 + 8 * 256 uint64, slot = 1 + event * 256 + tile index of CTA 0 */
int cv_gemm_set_trace(void* device_buffer);

/* Block-diagonal flash attention on tcgen05 (head_dim 96): tokens are window-major; query row i (window i / Wq)
 * attends the Wkv keys of the same window.  Covers Hiera's windowed, Q-pooled (Wq = Wkv/4) and global
 * (Wq = Wkv = tokens per image) attention (sam2 package, called from sam2_infer.py:226).  q/k/v are bf16 matrices
 * with row pitch ld* whose columns [col0 + h*96, col0 + (h+1)*96) hold head h; out is [Mq, heads*96] bf16.     */
int cv_attention_bf16(const void* q, long long ldq, int qcols, int qcol0, const void* k, long long ldk, int kcols,
                      int kcol0, const void* v, long long ldv, int vcols, int vcol0, int Mq, int Mkv, int Wq, int Wkv,
                      int heads, int head_dim, float scale, void* out, long long ld_out, void* stream);

/* ------------------------------------------------------------------ SAM 2.1 image path
 * replaces sam2_infer.py:220-275 (SAM2ImageWrapper.forward: sam2 image_encoder, conv_s0/conv_s1, sam_mask_decoder with
 * the wrapper's learned prompts, bilinear x4, MultiKernelRefinement :177-189), SAM2Transforms.__call__ (:49-51) and
 * postprocess_masks (:88-128) + the threshold / extent tail of circuit_analyzer.py:355-370.
 * The engine owns the device copies of the (load-time folded) weights and its activation workspace.           */
typedef struct cv_sam2_cfg {
  int32_t embed_dim, num_heads;  /* stage-1 width and heads (tiny/small 96/1)                        */
  int32_t stages[4];             /* blocks per stage                                                   */
  int32_t window_spec[4];        /* window size per stage                                              */
  int32_t global_blocks[8];      /* indices of global-attention blocks                                 */
  int32_t n_global;
  int32_t use_refinement;        /* wrapper built with use_refinement=True (sam2_infer.py:210)         */
  int32_t max_batch;             /* activation workspace is sized for this many images per call        */
  int32_t operand_fp16;          /* 16-bit tensor-core operand format: 0 = bf16, 1 = IEEE fp16 (saturating);
                                    the weights passed to cv_sam2_set_tensor(dtype 1) must use the same format */
} cv_sam2_cfg;
typedef struct cv_sam2 cv_sam2;

int cv_sam2_create(const cv_sam2_cfg* cfg, int device, cv_sam2** out);
int cv_sam2_destroy(cv_sam2* h);
/* Upload one named weight tensor from HOST memory (dtype 0 = float32, 1 = bfloat16 bits).  The names and layouts
 * are those produced by circuitvision_b200/sam2_weights.py (fold_state_dict) and listed in DESIGN.md.          */
int cv_sam2_set_tensor(cv_sam2* h, const char* name, const void* host_data, int dtype, long long numel);
/* Checks that every tensor the configuration needs is present and allocates the workspace. */
int cv_sam2_finalize(cv_sam2* h);
/* Re-sizes the activation workspace for a different max_batch (weights stay resident). */
int cv_sam2_set_max_batch(cv_sam2* h, int max_batch);
/* images (device): input_kind 0 = uint8 HWC [B,1024,1024,3] (ToTensor + Normalize fused; swap_rb applies the
 * channel swap of circuit_analyzer.py:343), 1 = float32 CHW [B,3,1024,1024] already normalised.
 * outputs (device, each nullable): low_res [B,1,256,256] f32, iou [B] f32, high_res [B,1,1024,1024] f32 (after the
 * refinement head), mask_u8 [B,out_h,out_w] = (postprocess_masks(high_res, (out_h,out_w)) > 0) * 255,
 * out_logits [B,out_h,out_w] f32 = postprocess_masks(high_res), extents [B,4] int32 (min x, min y, max x, max y of
 * the foreground; max < 0 when empty).                                                                        */
int cv_sam2_forward(cv_sam2* h, const void* images, int input_kind, int swap_rb, int B, float* low_res, float* iou,
                    float* high_res, uint8_t* mask_u8, int out_h, int out_w, float* out_logits, int* extents,
                    void* stream);
int cv_sam2_last_launches(cv_sam2* h);
/* Debug tap: with count_fp16_saturation != 0 every 16-bit activation buffer is scanned after the kernel that wrote it and
 * the number of entries that saturated in the fp32 -> fp16 conversion (+-65504) accumulates in the uint32 buffer
 * "satcount" (cv_sam2_read_buffer), reset at the start of each cv_sam2_forward.  Off by default (extra launches).  */
int cv_sam2_set_debug(cv_sam2* h, int count_fp16_saturation);
/* Parity taps: copy an internal activation buffer (DESIGN.md names them) to caller device memory. */
int cv_sam2_read_buffer(cv_sam2* h, const char* name, void* dst_device, long long bytes, void* stream);
/* SAM2Transforms.__call__ for one uint8 HWC image of any size -> float32 CHW [3,1024,1024]; tmp: H*1024*3 floats. */
int cv_sam2_preprocess(const uint8_t* img_hwc, int H, int W, int swap_rb, float* tmp, float* out_chw, void* stream);
/* Batched SAM2Transforms.__call__ on crop windows of whole pages (analysis_pipeline.py:177-208: crop_image_and_adjust_bboxes,
 * then segment_with_sam2 on the crop): image b = window [x0,x1) x [y0,y1) of page b, a uint8 HWC page of width pw at
 * pages + off.  geom: B device records of 32 bytes {int64 off; int32 pw, x0, y0, x1, y1, pad}.  tmp: B*max_crop_h*1024*3 floats.
 * out_chw: [B,3,1024,1024] f32, what cv_sam2_forward(input_kind 1) consumes.  Only the raw uint8 pages cross PCIe.          */
int cv_sam2_preprocess_pages(const uint8_t* pages, const void* geom, int B, int max_crop_h, int swap_rb, float* tmp,
                             float* out_chw, void* stream);
/* SAM2Transforms.postprocess_masks with hole/sprinkle filters off: bilinear (align_corners=False) resize of
 * [B,1,S,S] f32 logits to [B,1,H,W]; optional threshold mask + extents as in cv_sam2_forward.                 */
int cv_sam2_resize_logits(const float* logits, int B, int S, int H, int W, float* out_logits, uint8_t* mask_u8,
                          int* extents, void* stream);
/* MultiKernelRefinement.forward (sam2_infer.py:177-189) on [B,1,1024,1024] f32: branch weights w[j] = [4,k_j,k_j]
 * for k = 3,5,7,11 and b[j] = [4] (device), combiner weight cw[16] (device), bias cb.                         */
int cv_sam2_refine(const float* x, int B, const float* const* w, const float* const* b, const float* cw, float cb,
                   float* out, void* stream);

/* Per-kernel device timing for bench.py's roofline line: while enabled, every kernel the library launches is
 * bracketed by two CUDA events on its own stream.  cv_profile_count() waits for the recorded kernels and returns
 * the number of distinct kernel names; cv_profile_get(i, ...) returns name, launches, summed milliseconds and
 * summed algorithmic work (bytes for HBM-bound kernels, flops for tensor kernels; 0 when not annotated). */
int cv_profile_enable(int on);
int cv_profile_reset(void);
int cv_profile_count(void);
int cv_profile_get(int i, char* name, int name_cap, long long* launches, double* total_ms, double* work);

#ifdef __cplusplus
}
#endif
#endif /* CV_B200_H */
