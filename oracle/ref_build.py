#!/usr/bin/env python3
"""Builds oracle/_ref/libtaa_ref.so from the REFERENCE's own shader sources (TEST INFRASTRUCTURE).

  python3 oracle/ref_build.py --ref /root/reference --out oracle/_ref

The reference's device code is GLSL; neither glslang nor a Vulkan driver exists in the build container,
so the shader text is rewritten mechanically into C++ against oracle/glsl_shim.h and compiled with g++ -ffp-contract=off.
The sources are read where they lie under --ref; only generated files go to --out (git-ignored, shipped to the GPU box).
The rewrite rules are purely lexical and listed in glsl_shim.h; the algorithm is never restated here. tests/test_oracle_vs_ref.py
then checks the hand-written restatement (taa_oracle.cpp) against this library bit for bit.
"""
import argparse
import os
import re
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))

SWIZZLE = re.compile(r"\.([xyzw]{2,4}|[rgba]{2,4})\b")
FLOAT_LIT = re.compile(r"(?<![\w.])((?:\d+\.\d*|\.\d+)(?:[eE][-+]?\d+)?|\d+[eE][-+]?\d+)(?![\w.])")
RESOURCE = re.compile(r"^layout\s*\([^)]*\)\s*((?:(?:uniform|readonly|writeonly|restrict)\s+)+)(\w+)\s+(\w+)\s*;\s*$")
BLOCK_OPEN = re.compile(r"^layout\s*\(([^)]*)\)\s*uniform\s+(\w+)\s*\{\s*$")
GLOBAL_VAR = re.compile(r"^(Parameters|ivec2|vec2|vec3|vec4|float|int|uint|bool)\s+(\w+)(\s*=\s*[^;]+)?;\s*$")


def strip_comments(src: str) -> str:
    src = re.sub(r"/\*.*?\*/", lambda m: "\n" * m.group(0).count("\n"), src, flags=re.S)
    return re.sub(r"//[^\n]*", "", src)


def rewrite_params(line: str) -> str:
    """in/out/inout qualifiers of function parameters -> by value / by reference."""
    line = re.sub(r"\b(?:out|inout)\s+(\w+)\s+(\w+)", r"\1& \2", line)
    return re.sub(r"\bin\s+(\w+)\s+(\w+)", r"\1 \2", line)


# headers whose text holds fp arithmetic: spliced into the shader before rewriting, so that their literals become fp32 like the shader's
# (shader_cpu_common.h holds only #defines and stays a C preprocessor include)
INLINE_INCLUDES = ("Fxaa3_11_mod.h",)


def splice_includes(src: str, folder: str) -> str:
    def sub(m):
        return strip_comments(open(os.path.join(folder, m.group(1))).read()) if m.group(1) in INLINE_INCLUDES else m.group(0)
    return re.sub(r'^[ \t]*#include\s+"([^"]+)"[^\n]*$', sub, src, flags=re.M)


# ---- sharpen_cas.comp: its two includes are AMD's portability header ffx_a.h (every GLSL/HLSL/CPU type and intrinsic wrapper it has:
# ~1900 lines, most of them needing built-ins the shim does not have) and ffx_cas.h. What the shader USES of ffx_a.h is spliced in by
# name, read from the reference's file at build time; of ffx_cas.h, the non-packed GPU section (the CPU-side CasSetup and the fp16
# variant in the same file are not part of this shader's code path).
FFX_A_DEFINES = re.compile(r"^\s*#define\s+(A[PFU][1-4]|ASU[1-4]|AF[1-4]_AU[1-4]\(x\)|AU[1-4]_AF[1-4]\(x\)|AF[1-4]_\(a\)|AU[1-4]_\(a\))\s")
FFX_A_GLSL_FUNCS = re.compile(r"^\s*A[FU][1-4]\s+(AF[1-4]_x|AU[1-4]_x|AMax3F1|AMin3F1|ARcpF1|ASatF1|ABfe|ABfiM)\(")
FFX_A_COMMON_FUNCS = re.compile(r"^\s*A[FU][1-4]\s+(APrxLoSqrtF1|APrxLoRcpF1|APrxMedRcpF1|ARmp8x8)\(")


def ffx_a_subset(path: str) -> str:
    lines = strip_comments(open(path).read()).split("\n")
    a = next(i for i, l in enumerate(lines) if re.match(r"^#if defined\(A_GLSL\) && defined\(A_GPU\)", l))
    b = next(i for i, l in enumerate(lines) if re.match(r"^#if defined\(A_HLSL\) && defined\(A_GPU\)", l))
    out = [l for l in lines[a:b] if FFX_A_DEFINES.match(l) or FFX_A_GLSL_FUNCS.match(l)]
    out += [l for l in lines[b:] if FFX_A_COMMON_FUNCS.match(l)]
    return "\n".join(out)


def ffx_cas_gpu_section(path: str) -> str:
    raw = open(path).read()
    start = raw.index("#ifdef A_GPU", raw.index("NON-PACKED VERSION"))
    end = raw.index("#if defined(A_GPU) && defined(A_HALF)")
    return strip_comments(raw[start:end])


def cas_source(path: str) -> str:
    folder = os.path.dirname(path)
    src = strip_comments(open(path).read())
    src = re.sub(r'^[ \t]*#include\s+"ffx_a.h"[^\n]*$', lambda m: ffx_a_subset(os.path.join(folder, "ffx_a.h")), src, flags=re.M)
    src = re.sub(r'^[ \t]*#include\s+"ffx_cas.h"[^\n]*$', lambda m: ffx_cas_gpu_section(os.path.join(folder, "ffx_cas.h")), src, flags=re.M)
    # the shader's only imageLoad calls read the rgba16f source: the shim's imageLoad is the r32ui one (taa.comp's seg-mask), imageLoadF the float one
    return src.replace("imageLoad(", "imageLoadF(")


def transpile(path: str, ns: str, prelude: str = ""):
    """Returns (C++ text of the shader inside namespace glsl::<ns>, names of globals that carry an initialiser)."""
    if os.path.basename(path) == "sharpen_cas.comp":
        src = cas_source(path)
    else:
        src = splice_includes(strip_comments(open(path).read()), os.path.dirname(path))
    out, resets = [], []
    in_block = None   # (kind, name) while inside a struct / uniform block
    depth = 0
    for raw in src.split("\n"):
        line = raw.rstrip()
        st = line.strip()
        if st.startswith("#version") or st.startswith("#extension"):
            continue
        if re.match(r"^layout\s*\(\s*local_size_x", st):
            continue
        if not st.startswith("#include"):
            line = FLOAT_LIT.sub(lambda m: m.group(1) + "f", line)
            line = SWIZZLE.sub(lambda m: "._" + m.group(1) + "()", line)
        st = line.strip()
        m = BLOCK_OPEN.match(st)
        if m:  # uniform block (UBO or push constants): a struct plus one thread_local instance, declared at the closing brace
            in_block = ("uniform", m.group(2))
            block_start = len(out)
            out.append(f"struct {m.group(2)} {{")
            continue
        if in_block is None and re.match(r"^struct\s+\w+\s*\{\s*$", st):
            in_block = ("struct", st.split()[1])
            out.append(line)
            continue
        if in_block is not None:
            mm = re.match(r"^\}\s*(\w+)?\s*;\s*$", st)
            if mm:
                if in_block[0] == "uniform" and mm.group(1) is None:
                    # a block without an instance name: its members are globals of the shader (function parameters may shadow them)
                    members = [l.strip() for l in out[block_start + 1:] if l.strip()]
                    del out[block_start:]
                    out += [f"static thread_local {l}" for l in members]
                else:
                    out.append("};")
                    if in_block[0] == "uniform":
                        out.append(f"static thread_local {in_block[1]} {mm.group(1)};")
                in_block = None
            else:
                out.append(re.sub(r"\bbool\b", "bool32", line))  # std140: bool is 4 bytes
            continue
        m = RESOURCE.match(st)
        if m:
            out.append(f"static thread_local {m.group(2)} {m.group(3)};")
            continue
        depth_before = depth
        depth += line.count("{") - line.count("}")
        if depth_before == 0:
            m = GLOBAL_VAR.match(st)
            if m and "(" not in st.split("=")[0]:
                out.append(f"static thread_local {m.group(1)} {m.group(2)}{m.group(3) or ''};")
                if m.group(3):
                    resets.append((m.group(2), m.group(3).lstrip(" =")))
                continue
            line = re.sub(r"\bvoid\s+main\s*\(\s*\)", "void shader_main()", line)
            line = rewrite_params(line)
        out.append(line)
    body = "\n".join(out)
    reset_fn = "static inline void reset_globals() {" + " ".join(f"{n} = {v};" for n, v in resets) + "}"
    return f"{prelude}namespace glsl {{ namespace {ns} {{\n{body}\n{reset_fn}\n", resets


HARNESS_HEAD = r'''// GENERATED by oracle/ref_build.py from the reference's shader sources. Do not commit.
#include "glsl_shim.h"
#include "../../include/taa_b200.h"
#ifdef _OPENMP
#include <omp.h>
#endif
static glsl::Image mk(const taa_image& s, int w, int h, glsl::Format f) {
	glsl::Image i; i.data = (unsigned char*)s.data; i.pitch = s.pitch_bytes; i.w = w; i.h = h; i.fmt = f; return i;
}
'''

HARNESS_TAA = r'''
static_assert(sizeof(Parameters) == 176 && sizeof(Matrices) == 544, "interface blocks must keep the std140 layout");
static void run(const taa_resolve_images* im, const TaaUniforms* u, int in_w, int in_h, int out_w, int out_h, int y0, int y1, int nthreads) {
#ifdef _OPENMP
	if (nthreads <= 0) nthreads = omp_get_max_threads();
#pragma omp parallel num_threads(nthreads)
#endif
	{
		memcpy((void*)&ubo, u, sizeof(Matrices));
		uCurrentFrame = mk(im->color, in_w, in_h, F_RGBA16F);
		uCurrentDepth = mk(im->depth, in_w, in_h, F_R32F);
		uCurrentVelocity = mk(im->velocity, in_w, in_h, F_RGBA16F);
		uCurrentUvNrm = mk(im->uvnrm, in_w, in_h, F_RGBA32F);
		uCurrentMaterial = mk(im->matid, in_w, in_h, F_R32UI);
		uPreviousMaterial = mk(im->prev_matid, in_w, in_h, F_R32UI);
		uHistoryFrame = mk(im->history_in, out_w, out_h, F_RGBA16F);
		uHistoryDepth = mk(im->history_depth, in_w, in_h, F_R32F);
		uResultScreen = mk(im->result, out_w, out_h, F_RGBA16F);
		uResultHistory = mk(im->history_out, out_w, out_h, F_RGBA16F);
		uSegMask = mk(im->segmask, out_w, out_h, F_R32UI);
		uPreviousSegMask = mk(im->prev_segmask, out_w, out_h, F_R32UI);
		uDebug = mk(im->debug, out_w, out_h, F_RGBA16F);
#ifdef _OPENMP
#pragma omp for schedule(dynamic, 4)
#endif
		for (int y = y0; y < y1; ++y)
			for (int x = 0; x < out_w; ++x) {
				gl_GlobalInvocationID = uvec3{(uint)x, (uint)y, 0u};
				reset_globals();
				shader_main();
			}
	}
}
}}  // namespaces
extern "C" __attribute__((visibility("default"))) int taa_ref_resolve(const taa_resolve_images* im, const TaaUniforms* u, int in_w, int in_h,
                                                                      int out_w, int out_h, int y0, int y1, int nthreads) {
	if (!im || !u) return -1;
	glsl::ref_taa::run(im, u, in_w, in_h, out_w, out_h, y0, y1, nthreads);
	return 0;
}
'''

HARNESS_SHARPEN = r'''
static void run(const taa_image* src, const taa_image* dst, int w, int h, float factor) {
#ifdef _OPENMP
#pragma omp parallel
#endif
	{
		uInputFrame = mk(*src, w, h, F_RGBA16F);
		uOutput = mk(*dst, w, h, F_RGBA16F);
		pushConstants.sharpeningFactor = factor;
#ifdef _OPENMP
#pragma omp for schedule(dynamic, 4)
#endif
		for (int y = 0; y < h; ++y)
			for (int x = 0; x < w; ++x) {
				gl_GlobalInvocationID = uvec3{(uint)x, (uint)y, 0u};
				shader_main();
			}
	}
}
}}
extern "C" __attribute__((visibility("default"))) int taa_ref_sharpen(const taa_image* src, const taa_image* dst, int w, int h, float factor) {
	glsl::ref_sharpen::run(src, dst, w, h, factor);
	return 0;
}
'''

HARNESS_POST = r'''
static_assert(sizeof(PushConstants) == 84, "push constants must keep their layout");
static void run(const taa_image* src, const taa_image* dbg, const taa_image* dst, int w, int h, const TaaPostProcessPush* pc) {
#ifdef _OPENMP
#pragma omp parallel
#endif
	{
		uInputFrame = mk(*src, w, h, F_RGBA16F);
		taa_image none = {nullptr, 0, 0, 0};
		uDebugFrame = mk(dbg ? *dbg : none, w, h, F_RGBA16F);
		uOutput = mk(*dst, w, h, F_RGBA16F);
		memcpy((void*)&pushConstants, pc, sizeof(PushConstants));
#ifdef _OPENMP
#pragma omp for schedule(dynamic, 4)
#endif
		for (int y = 0; y < h; ++y)
			for (int x = 0; x < w; ++x) {
				gl_GlobalInvocationID = uvec3{(uint)x, (uint)y, 0u};
				shader_main();
			}
	}
}
}}
extern "C" __attribute__((visibility("default"))) int taa_ref_post_process(const taa_image* src, const taa_image* dbg, const taa_image* dst, int w, int h,
                                                                           const TaaPostProcessPush* pc) {
	glsl::ref_post::run(src, dbg, dst, w, h, pc);
	return 0;
}
'''

HARNESS_FXAA_PREPARE = r'''
static void run(const taa_image* src, const taa_image* dst, int w, int h) {
#ifdef _OPENMP
#pragma omp parallel
#endif
	{
		uInput = mk(*src, w, h, F_RGBA16F);
		uOutput = mk(*dst, w, h, F_RGBA16F);
#ifdef _OPENMP
#pragma omp for schedule(dynamic, 4)
#endif
		for (int y = 0; y < h; ++y)
			for (int x = 0; x < w; ++x) {
				gl_GlobalInvocationID = uvec3{(uint)x, (uint)y, 0u};
				shader_main();
			}
	}
}
}}
extern "C" __attribute__((visibility("default"))) int taa_ref_fxaa_prepare(const taa_image* src, const taa_image* dst, int w, int h) {
	glsl::ref_fxaa_prepare::run(src, dst, w, h);
	return 0;
}
'''

# antialias_fxaa.comp is compiled twice: with GL_ARB_gpu_shader5 predefined (what glslang does -> FXAA_GATHER4_ALPHA 1) and without
HARNESS_FXAA = r'''
static_assert(sizeof(PushConstants) == 32, "push constants must keep their layout");
static void run(const taa_image* src, const taa_image* seg, const taa_image* dst, int w, int h, const TaaFxaaPush* pc) {
#ifdef _OPENMP
#pragma omp parallel
#endif
	{
		static thread_local Image srcImage;
		srcImage = mk(*src, w, h, F_RGBA16F);
		texInput.t = &srcImage;
		uOutput = mk(*dst, w, h, F_RGBA16F);
		taaSegMask = mk(*seg, w, h, F_R32UI);
		memcpy((void*)&pushConstants, pc, sizeof(PushConstants));
#ifdef _OPENMP
#pragma omp for schedule(dynamic, 4)
#endif
		for (int y = 0; y < h; ++y)
			for (int x = 0; x < w; ++x) {
				gl_GlobalInvocationID = uvec3{(uint)x, (uint)y, 0u};
				shader_main();
			}
	}
}
}}
extern "C" __attribute__((visibility("default"))) int FXAA_ENTRY(const taa_image* src, const taa_image* seg, const taa_image* dst, int w, int h, const TaaFxaaPush* pc) {
	glsl::FXAA_NS::run(src, seg, dst, w, h, pc);
	return 0;
}
'''

# sharpen_cas.comp: 64 threads per workgroup, each writes 4 pixels of a 16x16 block (sharpen_cas.comp:30-52); dispatched as
# ((w + 15) / 16, (h + 15) / 16) workgroups (taa.hpp:1134)
HARNESS_CAS = r'''
static void run(const taa_image* src, const taa_image* dst, int w, int h, const TaaCasPush* pc) {
#ifdef _OPENMP
#pragma omp parallel
#endif
	{
		imgSrc = mk(*src, w, h, F_RGBA16F);
		imgDst = mk(*dst, w, h, F_RGBA16F);
		const0 = uvec4(pc->const0[0], pc->const0[1], pc->const0[2], pc->const0[3]);
		const1 = uvec4(pc->const1[0], pc->const1[1], pc->const1[2], pc->const1[3]);
#ifdef _OPENMP
#pragma omp for schedule(dynamic, 1)
#endif
		for (int gy = 0; gy < (h + 15) / 16; ++gy)
			for (int gx = 0; gx < (w + 15) / 16; ++gx)
				for (int l = 0; l < 64; ++l) {
					gl_WorkGroupID = uvec3{(uint)gx, (uint)gy, 0u};
					gl_LocalInvocationID = uvec3{(uint)l, 0u, 0u};
					shader_main();
				}
	}
}
}}
extern "C" __attribute__((visibility("default"))) int taa_ref_sharpen_cas(const taa_image* src, const taa_image* dst, int w, int h, const TaaCasPush* pc) {
	glsl::ref_cas::run(src, dst, w, h, pc);
	return 0;
}
'''

# (shader, namespace, harness, text put in front of the namespace)
SHADERS = [("taa.comp", "ref_taa", HARNESS_TAA, ""), ("sharpen.comp", "ref_sharpen", HARNESS_SHARPEN, ""), ("sharpen_cas.comp", "ref_cas", HARNESS_CAS, ""),
           ("post_process.comp", "ref_post", HARNESS_POST, ""),
           ("antialias_fxaa_prepare.comp", "ref_fxaa_prepare", HARNESS_FXAA_PREPARE, ""),
           ("antialias_fxaa.comp", "ref_fxaa_gather", HARNESS_FXAA, "#define GL_ARB_gpu_shader5 1\n#define FXAA_ENTRY taa_ref_fxaa_gather4\n#define FXAA_NS ref_fxaa_gather\n"),
           ("antialias_fxaa.comp", "ref_fxaa_offset", HARNESS_FXAA, "#define FXAA_ENTRY taa_ref_fxaa_offset\n#define FXAA_NS ref_fxaa_offset\n")]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--ref", default="/root/reference")
    ap.add_argument("--out", default=os.path.join(HERE, "_ref"))
    ap.add_argument("--cxx", default="/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++")
    a = ap.parse_args()
    shaders = os.path.join(a.ref, "shaders")
    if not os.path.isdir(shaders):
        print(f"no reference tree at {a.ref}: keeping whatever {a.out} holds", file=sys.stderr)
        return 0
    os.makedirs(a.out, exist_ok=True)
    objs = []
    for fname, ns, harness, prelude in SHADERS:
        text, _ = transpile(os.path.join(shaders, fname), ns, prelude)
        cpp = os.path.join(a.out, ns + ".cpp")
        with open(cpp, "w") as f:
            f.write(HARNESS_HEAD + text + harness)
        obj = cpp[:-4] + ".o"
        cmd = [a.cxx, "-O2", "-std=c++17", "-fPIC", "-ffp-contract=off", "-fno-fast-math", "-fopenmp", "-fvisibility=hidden", "-I", HERE, "-I", shaders,
               "-c", cpp, "-o", obj]
        subprocess.check_call(cmd)
        objs.append(obj)
    so = os.path.join(a.out, "libtaa_ref.so")
    subprocess.check_call([a.cxx, "-shared", "-fopenmp", "-o", so] + objs)
    for o in objs:
        os.remove(o)
    print(f"built {so} from {shaders}/{{{', '.join(sorted(set(s[0] for s in SHADERS)))}}}")
    return 0


if __name__ == "__main__":
    sys.exit(main())
