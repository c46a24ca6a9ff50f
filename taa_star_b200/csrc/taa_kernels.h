// taa_kernels.h — host-callable launchers of the CUDA kernels (internal to the library).
#pragma once
#include <cuda_runtime.h>
#include "taa_device.cuh"

namespace taa {

// taa.comp, fully general, exact arithmetic (taa_resolve_generic.cu)
// allow_specialised: a call whose switches are the reference's default pattern (BASELINE configs[0]) may run on taa_resolve_defaults_kernel —
// the same arithmetic with those switches folded at compile time and the colour taps shared through a shared-memory tile (bit-identical)
cudaError_t launch_resolve_generic(const ResolveArgs& args, bool allow_specialised, cudaStream_t stream);
bool defaults_kernel_supports(const ResolveArgs& args);
// the same exact arithmetic on a device-side list of pixels (y * out_w + x), count read on the device
cudaError_t launch_resolve_fixup(const ResolveArgs& args, const unsigned int* list, const unsigned int* count, bool write_screen, int num_sms,
                                 cudaStream_t stream);

// The BASELINE configs 2-5 family of settings (taa_dispatch.cu). The tuned kernels append the pixels whose `rectified` bit (or anti-ghosting
// predicate) needs the exact arithmetic to fix_list / *fix_count and zero *fix_count_next.
bool tuned_supports(const ResolveArgs& args);
// 64-wide shared-memory tiles, long column strips and a history window that runs one row ahead (taa_resolve_strip.cu): the rejection variants
cudaError_t launch_resolve_strip(const ResolveArgs& args, unsigned int* fix_list, unsigned int* fix_count, unsigned int* fix_count_next,
                                 bool fixup_all, cudaStream_t stream);

// the same contract as a streaming kernel (taa_resolve_stream.cu): the default. One warp per 62-column strip walks down the rows with rolling
// neighbourhood sums and a sliding history window; raw rows arrive through a per-warp ring of 2-D TMA boxes. stream_supports(): the
// images can be described by tensor maps (16-byte aligned base and pitch) and the driver exports cuTensorMapEncodeTiled.
bool stream_supports(const ResolveArgs& args);
// args.epilogue != 0 (the sharpening pass evaluated in the resolve's epilogue, written to args.final_img) is admissible for this call
bool stream_epilogue_ok(const ResolveArgs& args, bool fixup_all);
// hints: a device buffer of stream_hint_bytes() zeroed bytes owned by the context (or nullptr), hint_phase: a counter that advances by one per call —
// the units that were slow in the previous call are started first (see "slow units first" in taa_resolve_stream.cu)
size_t stream_hint_bytes();
// Row bands without a per-frame collective ("PEER variants" in taa_resolve_stream.cu): side 0 = the band above, 1 = the band below.
struct StreamPeers {
	unsigned char* nb_hist[2];   // the neighbour's history buffer with the parity of args.history_out, mapped into this process (nullptr: no neighbour)
	long long nb_pitch[2];
	int nb_y0[2];                // global row in the neighbour's buffer row 0
	int nb_band_rows[2];         // output rows the neighbour resolves
	unsigned int* nb_flags[2];   // the neighbour's flag block (TAA_BAND_FLAG_WORDS words), mapped
	unsigned int* flags;         // this band's flag block
	int halo, q, wait;           // halo rows; parity of the history buffer being written; 0: the first frame of a sequence (nothing to wait for)
};
cudaError_t launch_resolve_stream(const ResolveArgs& args, unsigned int* fix_list, unsigned int* fix_count, unsigned int* fix_count_next,
                                  bool fixup_all, int num_sms, unsigned int* hints, int hint_phase, const StreamPeers* peers, cudaStream_t stream);

// follow-on passes, one kernel each as the reference dispatches them (taa_post.cu)
struct PostImg { Img src; Img debug; ImgW dst; int w, h; };
// avk::blit_image (nearest, whole image -> whole image): io.src is src_w x src_h, io.dst is io.w x io.h
cudaError_t launch_blit_nearest(const PostImg& io, int src_w, int src_h, cudaStream_t stream);
cudaError_t launch_sharpen(const PostImg& io, float sharpeningFactor, cudaStream_t stream);           // sharpen.comp
cudaError_t launch_cas(const PostImg& io, const TaaCasPush& pc, cudaStream_t stream);                 // sharpen_cas.comp
cudaError_t launch_post_process(const PostImg& io, const TaaPostProcessPush& pc, cudaStream_t stream); // post_process.comp
// [sharpen.comp | sharpen_cas.comp] -> post_process.comp in one pass: no intermediate image
cudaError_t launch_sharpen_post(const PostImg& io, int sharpener, float sharpeningFactor, const TaaCasPush& cas, const TaaPostProcessPush& pc, cudaStream_t stream);

// the FXAA branch (taa_fxaa.cu). io.debug carries the segmentation mask for launch_fxaa; prepared = the source's alpha already
// holds the luma (antialias_fxaa_prepare.comp ran), else it is computed on the fly (both dispatches in one launch)
cudaError_t launch_fxaa_prepare(const PostImg& io, cudaStream_t stream);
cudaError_t launch_fxaa(const PostImg& io, const TaaFxaaPush& pc, bool prepared, cudaStream_t stream);

}  // namespace taa
