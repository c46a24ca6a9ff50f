/*
 * taa_b200.h — C-ABI of the B200-native temporal anti-aliasing resolve.
 *
 * This is the drop-in boundary for ONE hot path of cg-tuwien/TAA-STAR: the compute dispatches that
 * `taa<CF>::render()` records per frame (reference source/taa.hpp:974-1192): taa.comp, then
 * optionally sharpen.comp | sharpen_cas.comp, then post_process.comp. Everything else of the
 * reference (rasteriser, ray tracer, Vulkan framework, GUI) stays where it is.
 *
 * Conventions
 *  - plain C, no C++/torch types; all images are RAW DEVICE POINTERS to linear row-major storage,
 *    texel (x,y) at base + (y - y0)*pitch_bytes + x*bytes_per_texel (y0 = global row held in row 0
 *    of the buffer; 0 unless the context resolves a row band of a larger frame).
 *  - storage formats are the reference's (shaders/shader_cpu_common.h:60-77):
 *      colour / velocity / history / result / debug : rgba16f  (8 B/texel)
 *      depth                                         : D32 float (4 B/texel)
 *      material id, segmentation mask                : r32ui    (4 B/texel)
 *      uv+normal                                     : rgba32f  (16 B/texel)
 *  - parameter blocks are byte-identical to the reference's std140 blocks (offsets asserted below),
 *    so a host that fills `uniforms_for_taa` (taa.hpp:113-129) can pass that memory unchanged.
 *  - every call returns an int status (0 = OK, <0 = error); nothing throws across the boundary.
 *  - all device work is enqueued asynchronously on the caller's stream (a `cudaStream_t` passed as
 *    `void*`); parameter blocks are copied at call time, like the per-frame host-coherent UBO of
 *    the reference (taa.hpp:660-662, 995).
 *  - a context is not thread-safe (the reference is single-threaded, composition.hpp:4).
 *  - there is no CPU fallback: without a CUDA device every compute entry point fails with
 *    TAA_E_CUDA.
 */
#ifndef TAA_B200_H
#define TAA_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(_WIN32)
#define TAA_API __declspec(dllexport)
#else
#define TAA_API __attribute__((visibility("default")))
#endif

#define TAA_B200_ABI_VERSION 2

/* ---- status codes (replace the reference's C++ exceptions, avk.cpp:5520 / main.cpp:5074) ---- */
enum {
	TAA_OK              =  0,
	TAA_E_INVALID_ARG   = -1,
	TAA_E_UNSUPPORTED   = -2,
	TAA_E_CUDA          = -3,
	TAA_E_NCCL          = -4,
	TAA_E_HALO_OVERFLOW = -5,  /* a history gather left the rows available to this band */
	TAA_E_PEER_TIMEOUT = -6    /* taa_band_peers: a neighbour's boundary rows did not arrive */
};

/* ---- TAA_RTFLAG_* (shaders/shader_cpu_common.h:31-40) ---- */
enum {
	TAA_RTFLAG_OUT = 0x0001, TAA_RTFLAG_DIS = 0x0002, TAA_RTFLAG_NRM = 0x0004, TAA_RTFLAG_DPT = 0x0008,
	TAA_RTFLAG_MID = 0x0010, TAA_RTFLAG_LUM = 0x0020, TAA_RTFLAG_CNT = 0x0040, TAA_RTFLAG_ALL = 0x0080,
	TAA_RTFLAG_FXD = 0x0100, TAA_RTFLAG_FXA = 0x0200
};

typedef uint32_t taa_bool32; /* VkBool32 / GLSL std140 bool */

/* struct Parameters — source/taa.hpp:30-77 == shaders/taa.comp:50-95 (std140, 176 B) */
typedef struct TaaParameters {
	float      mAlpha;                    /*   0 */
	int32_t    mColorClampingOrClipping;  /*   4  0 none, 1 clamp, 2 clip (fast), 3 clip (slow) */
	taa_bool32 mDepthCulling;             /*   8 */
	taa_bool32 mUnjitterNeighbourhood;    /*  12 */
	taa_bool32 mUnjitterCurrentSample;    /*  16 */
	float      mUnjitterFactor;           /*  20 */
	taa_bool32 mPassThrough;              /*  24 */
	taa_bool32 mUseYCoCg;                 /*  28 */
	taa_bool32 mShrinkChromaAxis;         /*  32 */
	taa_bool32 mVarianceClipping;         /*  36 */
	taa_bool32 mShapedNeighbourhood;      /*  40 */
	taa_bool32 mLumaWeightingLottes;      /*  44 */
	float      mVarClipGamma;             /*  48 */
	float      mMinAlpha;                 /*  52 */
	float      mMaxAlpha;                 /*  56 */
	float      mRejectionAlpha;           /*  60 */
	taa_bool32 mRejectOutside;            /*  64 */
	int32_t    mUseVelocityVectors;       /*  68  0 off, 1 movers only, 2 everything */
	int32_t    mVelocitySampleMode;       /*  72  0 simple, 1 3x3 "longest", 2 3x3 closest */
	int32_t    mInterpolationMode;        /*  76  0 bilinear, 1 b-spline, 2 catmull-rom */
	taa_bool32 mToneMapLumaKaris;         /*  80 */
	taa_bool32 mAddNoise;                 /*  84 */
	float      mNoiseFactor;              /*  88 */
	taa_bool32 mReduceBlendNearClamp;     /*  92 */
	taa_bool32 mDynamicAntiGhosting;      /*  96 */
	taa_bool32 mVelBasedAlpha;            /* 100 */
	float      mVelBasedAlphaMax;         /* 104 */
	float      mVelBasedAlphaFactor;      /* 108 */
	taa_bool32 mRayTraceAugment;          /* 112 */
	uint32_t   mRayTraceAugmentFlags;     /* 116 */
	float      mRayTraceAugment_WNrm;     /* 120 */
	float      mRayTraceAugment_WDpt;     /* 124 */
	float      mRayTraceAugment_WMId;     /* 128 */
	float      mRayTraceAugment_WLum;     /* 132 */
	float      mRayTraceAugment_Thresh;   /* 136 */
	int32_t    mRayTraceHistoryCount;     /* 140 */
	float      mDebugMask[4];             /* 144 */
	int32_t    mDebugMode;                /* 160 */
	float      mDebugScale;               /* 164 */
	taa_bool32 mDebugCenter;              /* 168 */
	taa_bool32 mDebugToScreenOutput;      /* 172 */
} TaaParameters;

/* struct uniforms_for_taa — source/taa.hpp:113-129 == `uniform Matrices`, shaders/taa.comp:102-118 (544 B) */
typedef struct TaaUniforms {
	float         mHistoryViewProjMatrix[16]; /*   0  column-major (glm::mat4) */
	float         mInverseViewProjMatrix[16]; /*  64 */
	TaaParameters param[2];                   /* 128, 304 */
	float         mJitterNdc[4];              /* 480  .xy used; NDC units */
	float         mSinTime[4];                /* 496 */
	taa_bool32    splitScreen;                /* 512 */
	int32_t       splitX;                     /* 516 */
	taa_bool32    mUpsampling;                /* 520 */
	taa_bool32    mBypassHistoryUpdate;       /* 524 */
	taa_bool32    mResetHistory;              /* 528 */
	float         mCamNearPlane;              /* 532 */
	float         mCamFarPlane;               /* 536 */
	float         pad1;                       /* 540 */
} TaaUniforms;

/* push_constants_for_sharpener — taa.hpp:84-86 / sharpen.comp:13-15 */
typedef struct TaaSharpenPush { float sharpeningFactor; } TaaSharpenPush;
/* push_constants_for_cas — taa.hpp:88-91 / sharpen_cas.comp:15-18 */
typedef struct TaaCasPush { uint32_t const0[4]; uint32_t const1[4]; } TaaCasPush;
/* push_constants_for_fxaa — taa.hpp:93-99 / antialias_fxaa.comp:14-20 */
typedef struct TaaFxaaPush {
	float fxaaQualityRcpFrame[2];
	float fxaaQualitySubpix;
	float fxaaQualityEdgeThreshold;
	float fxaaQualityEdgeThresholdMin;
	float pad1, pad2, pad3;
} TaaFxaaPush;
/* push_constants_for_postprocess — taa.hpp:101-111 / post_process.comp:13-23 (84 B) */
typedef struct TaaPostProcessPush {
	int32_t    zoomSrcLTWH[4];
	int32_t    zoomDstLTWH[4];
	float      debugL_mask[4];
	float      debugR_mask[4];
	taa_bool32 zoom;
	taa_bool32 showZoomBox;
	int32_t    splitX;
	taa_bool32 debugL_show;
	taa_bool32 debugR_show;
} TaaPostProcessPush;

#if defined(__cplusplus) || (defined(__STDC_VERSION__) && __STDC_VERSION__ >= 201112L)
#ifdef __cplusplus
#define TAA_STATIC_ASSERT(c, m) static_assert(c, m)
#else
#define TAA_STATIC_ASSERT(c, m) _Static_assert(c, m)
#endif
TAA_STATIC_ASSERT(sizeof(TaaParameters) == 176, "Parameters must be 176 B (std140, taa.hpp:78)");
TAA_STATIC_ASSERT(offsetof(TaaParameters, mVarClipGamma) == 48, "Parameters layout");
TAA_STATIC_ASSERT(offsetof(TaaParameters, mInterpolationMode) == 76, "Parameters layout");
TAA_STATIC_ASSERT(offsetof(TaaParameters, mRayTraceAugmentFlags) == 116, "Parameters layout");
TAA_STATIC_ASSERT(offsetof(TaaParameters, mDebugMask) == 144, "Parameters layout");
TAA_STATIC_ASSERT(offsetof(TaaParameters, mDebugToScreenOutput) == 172, "Parameters layout");
TAA_STATIC_ASSERT(sizeof(TaaUniforms) == 544, "uniforms_for_taa must be 544 B (taa.hpp:130)");
TAA_STATIC_ASSERT(offsetof(TaaUniforms, param) == 128, "uniforms layout");
TAA_STATIC_ASSERT(offsetof(TaaUniforms, mJitterNdc) == 480, "uniforms layout");
TAA_STATIC_ASSERT(offsetof(TaaUniforms, splitScreen) == 512, "uniforms layout");
TAA_STATIC_ASSERT(offsetof(TaaUniforms, mCamFarPlane) == 536, "uniforms layout");
TAA_STATIC_ASSERT(sizeof(TaaCasPush) == 32, "CAS push constants are 2 x uvec4");
TAA_STATIC_ASSERT(sizeof(TaaFxaaPush) == 32, "FXAA push constants are 32 B");
TAA_STATIC_ASSERT(sizeof(TaaPostProcessPush) == 84, "post-process push constants are 84 B");
#endif

/* ---- images ---- */
typedef struct taa_image {
	void*   data;         /* device pointer (host pointer only for the *_host entry points) */
	int64_t pitch_bytes;  /* >= width * bytes_per_texel, multiple of the texel size */
	int32_t y0;           /* global row index stored in buffer row 0 (band sharding); 0 for whole frames */
	int32_t rows;         /* rows present in the buffer */
} taa_image;

/*
 * The descriptor set of taa.comp (bindings at shaders/taa.comp:24-38; what render() binds to them:
 * source/taa.hpp:1009-1025). Optional images may have data == NULL.
 */
typedef struct taa_resolve_images {
	taa_image color;          /* b1  uCurrentFrame     rgba16f, input res            (required) */
	taa_image depth;          /* b2  uCurrentDepth     D32,     input res            (required) */
	taa_image velocity;       /* b7  uCurrentVelocity  rgba16f, input res            (required) */
	taa_image history_in;     /* b3  uHistoryFrame     rgba16f, output res           (required) */
	taa_image history_depth;  /* b4  uHistoryDepth     D32 = previous frame's depth  (if mDepthCulling) */
	taa_image history_out;    /* b8  uResultHistory    rgba16f, output res, written  (required) */
	taa_image result;         /* b5  uResultScreen     rgba16f, output res, written  (optional) */
	taa_image debug;          /* b6  uDebug            rgba16f, written              (optional; the reference always writes it) */
	taa_image segmask;        /* b9  uSegMask          r32ui, written                (if mRayTraceAugment) */
	taa_image prev_segmask;   /* b12 uPreviousSegMask  r32ui                         (if TAA_RTFLAG_CNT) */
	taa_image matid;          /* b10 uCurrentMaterial  r32ui                         (if mRayTraceAugment) */
	taa_image prev_matid;     /* b11 uPreviousMaterial r32ui                         (if TAA_RTFLAG_DIS) */
	taa_image uvnrm;          /* b13 uCurrentUvNrm     rgba32f                       (if TAA_RTFLAG_NRM) */
	taa_image mask;           /* --  r32ui, written: bit0 = rejected, bit1 = rectified, bits2-3 = clip mode
	                                 (taa.comp:739-740, 845; only visible through debug mode 2 in the reference) (optional) */
} taa_resolve_images;

/* the follow-on passes render() records after taa.comp (taa.hpp:1111-1159) */
typedef struct taa_post_chain {
	int32_t            sharpener;          /* mSharpener: 0 off, 1 sharpen.comp, 2 sharpen_cas.comp (taa.hpp:1418) */
	TaaSharpenPush     sharpen;            /* used if sharpener == 1 */
	TaaCasPush         cas;                /* used if sharpener == 2 (fill with taa_cas_setup) */
	int32_t            postprocess;        /* mPostProcessEnabled (taa.hpp:1352) */
	TaaPostProcessPush pp;                 /* used if postprocess != 0 */
	int32_t            fxaa;               /* run FXAA on the pixels the seg-mask marks (taa.hpp:1061: mRayTraceAugmentFlags & TAA_RTFLAG_FXA);
	                                          needs images->segmask; runs before the sharpener like in render() */
	TaaFxaaPush        fxaa_pc;            /* used if fxaa != 0 (fill with taa_fxaa_default) */
} taa_post_chain;

/* ---- context ---- */
enum {
	TAA_FLAG_DEFAULT     = 0,
	/* Arithmetic of the resolve kernels.
	 * default: settings blocks of the tuned family (YCoCg variance clip, clipAabb, Catmull-Rom history, velocity
	 *        reprojection; BASELINE configs 2-5) run on the shared-memory tiled kernel: coordinates and every
	 *        rejection predicate are evaluated exactly, colour filtering is re-associated and contracted (within
	 *        ~1e-5 of the exact result, gate 2^-10), and the pixels whose `rectified` predicate sits near its
	 *        threshold (or whose filtered history alpha cancels to ~0 under dynamic anti-ghosting) are recomputed
	 *        exactly by a second small launch, so integer masks are bit-exact. That second launch happens only when
	 *        it has something to decide: `mask` is bound (`rectified` is reported through it alone) or
	 *        mDynamicAntiGhosting is on (an undecided `rejected` changes the colour).
	 *        Every other settings block runs on the exact general kernel.
	 * EXACT: always the general kernel: every fp32 operation in the order of the reference shader with IEEE
	 *        round-to-nearest and no contraction; all outputs bit-identical to oracle/ on finite inputs. */
	TAA_FLAG_EXACT = 1u << 0,
	/* testing aid: the tuned kernel hands EVERY pixel to its exact fix-up pass (validates that pass bit for bit) */
	TAA_FLAG_FIXUP_ALL = 1u << 1
};

typedef struct taa_desc {
	uint32_t struct_size;    /* = sizeof(taa_desc) */
	uint32_t abi_version;    /* = TAA_B200_ABI_VERSION */
	int32_t  in_width;       /* input (lo-res) frame size: colour/depth/velocity, taa.comp:11 */
	int32_t  in_height;
	int32_t  out_width;      /* output (hi-res) frame size: history/result/debug, taa.comp:12 */
	int32_t  out_height;
	int32_t  band_y0;        /* first output row this context resolves (0 for whole frames) */
	int32_t  band_rows;      /* number of output rows this context resolves (out_height for whole frames) */
	int32_t  device;         /* CUDA device ordinal, -1 = current */
	uint32_t flags;          /* TAA_FLAG_* */
} taa_desc;

typedef struct taa_ctx taa_ctx;

TAA_API int         taa_abi_version(void);
TAA_API int         taa_create(taa_ctx** out_ctx, const taa_desc* desc);
TAA_API void        taa_destroy(taa_ctx* ctx);
TAA_API const char* taa_last_error_string(const taa_ctx* ctx); /* ctx may be NULL: last creation error */
TAA_API const char* taa_status_string(int status);

/*
 * taa_resolve — the call that replaces `dispatch(taa.comp)` (taa.hpp:1008-1027).
 * Whole-frame contexts only; images are tightly pitched (pitch = width * texel size).
 * `params` points to a TaaUniforms block. The screen result is not written (history only);
 * use taa_resolve_ex for all outputs.
 */
TAA_API int taa_resolve(taa_ctx* ctx, const void* color, const void* depth, const void* motion,
                        const void* history_in, void* history_out, const TaaUniforms* params, void* stream);

/* All bindings of taa.comp. */
TAA_API int taa_resolve_ex(taa_ctx* ctx, const taa_resolve_images* images, const TaaUniforms* params, void* stream);

/*
 * taa_frame — taa.comp plus the follow-on passes of render() (taa.hpp:1111-1159) in one call:
 * resolve -> [FXAA] -> [sharpen | CAS] -> [post-process]; `final` receives the image render() would blit to
 * the swapchain (taa.hpp:1161-1169). images->result may be NULL when a sharpener or post-process
 * consumes it. With a sharpener and an identity post-process (no zoom box, no splitter, no debug view) on the tuned
 * settings family without a mask binding the whole chain is ONE launch: the sharpening pass is evaluated in the
 * streaming resolve's epilogue on the resolved rows it still holds in registers, the unsharpened image is never
 * written. Otherwise the resolve writes it (to images->result or an internal scratch image) and one follow-on
 * launch evaluates [sharpen | CAS] + post-process together.
 */
TAA_API int taa_frame(taa_ctx* ctx, const taa_resolve_images* images, const TaaUniforms* params,
                      const taa_post_chain* chain, const taa_image* final, void* stream);

/* the follow-on passes on their own (unfused; what the reference dispatches one by one) */
TAA_API int taa_sharpen(taa_ctx* ctx, const taa_image* src, const taa_image* dst, const TaaSharpenPush* pc, void* stream);       /* sharpen.comp, taa.hpp:1117-1125 */
TAA_API int taa_sharpen_cas(taa_ctx* ctx, const taa_image* src, const taa_image* dst, const TaaCasPush* pc, void* stream);       /* sharpen_cas.comp, taa.hpp:1126-1135 */
TAA_API int taa_post_process(taa_ctx* ctx, const taa_image* src, const taa_image* debug, const taa_image* dst,
                             const TaaPostProcessPush* pc, void* stream);                                                       /* post_process.comp, taa.hpp:1141-1159 */

/* the FXAA branch (taa.hpp:1061-1107). `segmask` is the r32ui image taa.comp wrote (bits 0-1 == 1 marks a pixel for FXAA). */
TAA_API int taa_fxaa_prepare(taa_ctx* ctx, const taa_image* src, const taa_image* dst, void* stream);                            /* antialias_fxaa_prepare.comp, taa.hpp:1068-1075 */
TAA_API int taa_fxaa(taa_ctx* ctx, const taa_image* src_prepared, const taa_image* segmask, const taa_image* dst,
                     const TaaFxaaPush* pc, void* stream);                                                                      /* antialias_fxaa.comp, taa.hpp:1086-1094 */
/* both dispatches in one launch: reads the screen result itself, bit-identical to taa_fxaa_prepare + taa_fxaa */
TAA_API int taa_fxaa_fused(taa_ctx* ctx, const taa_image* src, const taa_image* segmask, const taa_image* dst,
                           const TaaFxaaPush* pc, void* stream);

/*
 * Row-band sharding without a per-frame collective (new work: the reference is single-GPU, main.cpp:4973; BASELINE configs[3]).
 * A band context whose neighbours' history buffers are mapped into this process stores the first / last `halo_rows` rows of
 * history_out a second time, straight into the neighbours' halo rows (peer memory over NVLink), from the resolve kernel itself,
 * and signals a counter in the neighbour's flag block; the neighbour's boundary units of the NEXT frame wait on that counter
 * before they read the halo. One launch per band and frame, no send/recv, no host synchronisation between the ranks.
 *   - every rank allocates its two history buffers and one flag block of TAA_BAND_FLAG_WORDS uint32 (taa_device_alloc),
 *     exports them (taa_ipc_export), exchanges the handles once (any transport) and opens its neighbours' (taa_ipc_open);
 *   - taa_band_peers() registers them and zeroes the own flag block: all ranks must have returned from it before any of them
 *     resolves (one barrier at set-up), and all ranks must write the SAME ping-pong index (history[k]) in the same frame;
 *   - calls on such a context must be served by the streaming kernel alone (config 2 family without a mask binding and
 *     without mDynamicAntiGhosting: the exact fix-up pass would rewrite pixels a neighbour already holds) — other calls fail;
 *   - a neighbour that never signals is reported by taa_poll_status as TAA_E_PEER_TIMEOUT after ~17 s, not as a hang.
 * up / down may be NULL (first / last band).
 */
#define TAA_BAND_FLAG_WORDS 16
typedef struct taa_band_peer {
	void*     history[2];  /* the neighbour's history ping-pong buffers as mapped into this process; [k] pairs with own history k */
	int64_t   row_pitch;
	int32_t   y0;          /* global row index stored in the neighbour's buffer row 0 */
	int32_t   band_rows;   /* output rows the neighbour resolves per frame */
	uint32_t* flags;       /* the neighbour's flag block as mapped into this process */
} taa_band_peer;
TAA_API int taa_band_peers(taa_ctx* ctx, const taa_band_peer* up, const taa_band_peer* down, void* own_history0, void* own_history1,
                           uint32_t* own_flags, int32_t halo_rows);
/* device memory that can be exported to other processes of the node (plain cudaMalloc: not from a pooled allocator) */
TAA_API void* taa_device_alloc(size_t bytes);
TAA_API void  taa_device_free(void* p);
#define TAA_IPC_HANDLE_BYTES 64
TAA_API int taa_ipc_export(const void* device_alloc_ptr, void* handle_out /* TAA_IPC_HANDLE_BYTES */);
TAA_API int taa_ipc_open(const void* handle /* TAA_IPC_HANDLE_BYTES */, void** mapped_out);
TAA_API int taa_ipc_close(void* mapped);

/* number of CUDA kernels launched through this context so far (bench.py reports it as gpu_launches) */
TAA_API long long taa_launch_count(const taa_ctx* ctx);

/* pixels the last resolve of this context handed from the tuned kernel to its exact fix-up pass (0 when the general
 * kernel ran); synchronises `stream`. Diagnostic only. */
TAA_API long long taa_fixup_pixels(taa_ctx* ctx, void* stream);

/* reads and clears the device-side status word of the last resolves (halo overflow); synchronises `stream` */
TAA_API int taa_poll_status(taa_ctx* ctx, void* stream);

/* ---- host-side helpers that mirror taa.hpp (pure CPU, no device needed) ---- */

/* CasSetup as called at taa.hpp:965 (shaders/ffx_cas.h:375-394): sharpen-only, in size == out size */
TAA_API void taa_cas_setup(TaaCasPush* out, float sharpness, float out_width, float out_height);

/* Parameters defaults of taa.hpp:31-76 */
TAA_API void taa_parameters_default(TaaParameters* out);
/* uniforms defaults: identity matrices, both param sets default, everything else zero */
TAA_API void taa_uniforms_default(TaaUniforms* out);
/* push_constants_for_fxaa defaults (taa.hpp:93-99) with fxaaQualityRcpFrame = 1 / (w, h) as update() sets it (taa.hpp:953) */
TAA_API void taa_fxaa_default(TaaFxaaPush* out, int32_t w, int32_t h);
/* push_constants_for_postprocess defaults as initialised at taa.hpp:352-359 for a w x h target */
TAA_API void taa_postprocess_default(TaaPostProcessPush* out, int32_t w, int32_t h);

/*
 * get_jitter_offset_for_frame (taa.hpp:150-233): sample_distribution 0 circular quad, 1 uniform4 helix,
 * 2 Halton(2,3)x8, 3 Halton(2,3)x16, 4 regular 16, 5 custom (debug offsets in pixel units).
 * Returns the pattern length (>0) or a negative status. out_ndc = offset in NDC units (pixel offset * 2/res).
 */
typedef struct taa_jitter_settings {
	int32_t mSampleDistribution;   /* taa.hpp:1353 (default 1) */
	int32_t mFixedJitterIndex;     /* taa.hpp:1404 (default -1) */
	float   mJitterExtraScale;     /* taa.hpp:1405 (default 1) */
	int32_t mJitterSlowMotion;     /* taa.hpp:1406 (default 1) */
	float   mJitterRotateDegrees;  /* taa.hpp:1407 (default 0) */
	const float* mDebugSampleOffsets; /* taa.hpp:1424: n x vec2, pixel units (pattern 5) */
	int32_t mDebugSampleOffsetsCount;
} taa_jitter_settings;
TAA_API int taa_jitter_offset_for_frame(const taa_jitter_settings* s, int32_t in_width, int32_t in_height,
                                        int64_t frame_id, float out_ndc[2]);
/* helpers::halton (source/helper_functions.hpp:9-17) */
TAA_API float taa_halton(int32_t index, int32_t base);
/* get_jittered_projection_matrix (taa.hpp:243-253): out = translate(jx, jy, 0) * proj, column-major */
TAA_API void taa_jittered_projection(const float proj[16], float jx, float jy, float out[16]);
/* the two matrices render() uploads (taa.hpp:993-994): inverse(P_cur * V_cur), P_prev * V_prev */
TAA_API int  taa_reprojection_matrices(const float proj_cur[16], const float view_cur[16],
                                       const float proj_prev[16], const float view_prev[16],
                                       float out_inverse_view_proj[16], float out_history_view_proj[16]);

/*
 * ---- the invokee: a C handle for `class taa<CF>` (taa.hpp:26-1427) ----
 * Owns what the reference's class owns: result / history / temp / debug / post-process / seg-mask
 * images x CF (taa.hpp:294-340), the per-frame uniforms, the history ring indexing
 * (`last = (i + CF - 1) % CF`, taa.hpp:980-981) and the jitter sequence. The caller keeps owning
 * the G-buffers and only registers pointers (set_source_image_views, taa.hpp:263-360).
 */
typedef struct taa_invokee taa_invokee;

typedef struct taa_source_views {          /* one frame-in-flight slot of set_source_image_views */
	const void* color;      /* rgba16f in_w x in_h, tightly pitched */
	const void* depth;      /* D32 */
	const void* uvnrm;      /* rgba32f (may be NULL) */
	const void* velocity;   /* rgba16f */
	const void* matid;      /* r32ui (may be NULL) */
	const void* raytraced;  /* rgba16f (may be NULL; unused: the ray tracer is out of scope) */
} taa_source_views;

typedef struct taa_invokee_settings {      /* the non-shader members of class taa (taa.hpp:1350-1424) */
	taa_bool32 mTaaEnabled;            /* 1351 */
	taa_bool32 mPostProcessEnabled;    /* 1352 */
	taa_bool32 mResetHistory;          /* 1354: consumed by the next update() */
	taa_bool32 mSplitScreen;           /* 1411 */
	int32_t    mSplitX;                /* 1412 */
	int32_t    mSharpener;             /* 1418 */
	float      mSharpenFactor;         /* 1419 */
	taa_bool32 mResetHistoryOnChange;  /* 1421 */
	taa_jitter_settings jitter;
} taa_invokee_settings;

TAA_API int  taa_invokee_create(taa_invokee** out, int32_t concurrent_frames, int32_t device, uint32_t flags);
TAA_API void taa_invokee_destroy(taa_invokee* t);
TAA_API const char* taa_invokee_last_error(const taa_invokee* t);
/* set_source_image_views (taa.hpp:263): views[CF]; allocates the owned images at target resolution */
TAA_API int  taa_invokee_set_source_image_views(taa_invokee* t, int32_t target_w, int32_t target_h,
                                                int32_t in_w, int32_t in_h, const taa_source_views* views);
TAA_API TaaParameters*        taa_invokee_parameters(taa_invokee* t, int32_t index /*0|1*/);   /* mParameters */
TAA_API taa_invokee_settings* taa_invokee_settings_ptr(taa_invokee* t);
TAA_API TaaPostProcessPush*   taa_invokee_postprocess(taa_invokee* t);
/* get_jittered_projection_matrix + save_history_proj_matrix (taa.hpp:235-259); frame_id selects slot frame_id % CF */
TAA_API int  taa_invokee_get_jittered_projection_matrix(taa_invokee* t, const float proj[16], int64_t frame_id,
                                                        float out_proj[16], float out_jitter_ndc[2]);
TAA_API int  taa_invokee_save_history_proj_matrix(taa_invokee* t, const float proj[16], int64_t frame_id);
/* update() (taa.hpp:894-971): view matrix, time, near/far come from the camera the reference queries */
TAA_API int  taa_invokee_update(taa_invokee* t, int64_t frame_id, const float view[16], float time_s,
                                float cam_near, float cam_far);
/* render() (taa.hpp:974-1192): enqueues the frame on `stream`; returns the image render() would blit */
TAA_API int  taa_invokee_render(taa_invokee* t, int64_t frame_id, void* stream, const void** out_final_image);
/* duration() (taa.hpp:367-374): device time of the last completed render() in ms (CUDA events) */
TAA_API float taa_invokee_duration(taa_invokee* t);
/* images owned by the invokee, for inspection (slot = frame_id % CF) */
enum { TAA_IMG_RESULT = 0, TAA_IMG_HISTORY = 1, TAA_IMG_DEBUG = 2, TAA_IMG_POSTPROCESS = 3, TAA_IMG_SEGMASK = 4, TAA_IMG_TEMP0 = 5, TAA_IMG_TEMP1 = 6 };
TAA_API void* taa_invokee_image(taa_invokee* t, int32_t which, int32_t slot);
/* kernels launched by this invokee so far */
TAA_API long long taa_invokee_launch_count(const taa_invokee* t);
/* bytes taa_invokee_frame_host has copied host -> device so far. Depth (4 of the 20 bytes per pixel) is uploaded only when a dispatch of the
 * frame can read it (depth culling, matrix reprojection, closest-depth velocity, segmentation mask); bench.py reports the per-frame delta. */
TAA_API long long taa_invokee_h2d_bytes(const taa_invokee* t);
/* the uniforms the last update()/render() produced (for parity tests against the oracle) */
TAA_API const TaaUniforms* taa_invokee_uniforms(const taa_invokee* t);

/*
 * Host-buffer frame: what a Vulkan host without CUDA interop would call. Copies this frame's G-buffer
 * from (pinned) host memory, runs render(), copies the final image back. Copies and compute of up to
 * CF frames overlap on internal streams, like the reference's CF frames in flight (main.cpp:341).
 * `out_final_host` receives out_w*out_h*8 bytes; it is complete after taa_invokee_wait(frame_id).
 */
TAA_API int  taa_invokee_frame_host(taa_invokee* t, int64_t frame_id, const taa_source_views* host_views,
                                    const float view[16], const float proj[16], float time_s,
                                    float cam_near, float cam_far, void* out_final_host);
TAA_API int  taa_invokee_wait(taa_invokee* t, int64_t frame_id);
/*
 * ---- settings files: writeSettingsToIni / readSettingsFromIni, taa.hpp:1198-1339 ----
 * The INI text of the reference's mINI dependency (external/include/mini/ini.h): sections TAA_Param_0, TAA_Param_1, TAA_Primary,
 * TAA_Postprocess with the reference's key names (names are not case sensitive), value formats of source/IniUtil.cpp:52-56,104-109.
 * An absent or empty value keeps the current setting. Pure host code.
 * write: returns the size of the text including the terminating NUL (call with out = NULL to size the buffer), < 0 on error.
 * read : `offsets` receives mDebugSampleOffsets (the current ones resized to max(1, size) vec2s, then the keys that are present) and
 *        becomes s->jitter.mDebugSampleOffsets; offsets_cap counts vec2s.
 *        A value that is not a number (where the reference's std::stoul / stol / stof would throw) gives TAA_E_INVALID_ARG; the valid
 *        keys are applied all the same; taa_settings_ini_last_error() names the first offending key (thread-local).
 */
TAA_API int32_t taa_settings_write_ini(const TaaParameters params[2], const taa_invokee_settings* s, const TaaPostProcessPush* pp,
                                       char* out, int32_t cap);
TAA_API int  taa_settings_read_ini(const char* text, TaaParameters params[2], taa_invokee_settings* s, TaaPostProcessPush* pp,
                                   float* offsets, int32_t offsets_cap);
TAA_API const char* taa_settings_ini_last_error(void);
/* the same on the invokee's own mParameters / settings / mPostProcessPushConstants (it owns the sample-offset storage) */
TAA_API int32_t taa_invokee_write_settings_ini(taa_invokee* t, char* out, int32_t cap);
TAA_API int  taa_invokee_read_settings_ini(taa_invokee* t, const char* text);

/* pinned host allocation helpers for the host-buffer path */
TAA_API void* taa_host_alloc(size_t bytes);
TAA_API void  taa_host_free(void* p);

#ifdef __cplusplus
}
#endif
#endif /* TAA_B200_H */
