"""Generates the settings-file fixtures from REFERENCE code compiled as-is (run in the build container only):

  taa_settings_written.ini   what taa<CF>::writeSettingsToIni + mINI generate() produce for the values of taa_settings_values.json
  taa_settings_read.json     what taa<CF>::readSettingsFromIni makes of taa_settings_input.ini (hand-written: mixed case, comments,
                             missing and empty keys, "true", padded values), starting from those same values

The bodies of writeSettingsToIni / readSettingsFromIni are extracted from source/taa.hpp at run time and compiled inside a stub class
that has the members they touch; source/IniUtil.cpp and external/include/mini/ini.h are compiled from where they lie. Nothing of the
reference is copied into the repository: only the produced text and numbers are.
"""
import json
import os
import re
import subprocess
import sys
import tempfile

REF = sys.argv[1] if len(sys.argv) > 1 else "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))

B32 = ["mPassThrough", "mShapedNeighbourhood", "mVarianceClipping", "mUseYCoCg", "mLumaWeightingLottes", "mDepthCulling", "mRejectOutside",
       "mUnjitterNeighbourhood", "mUnjitterCurrentSample", "mToneMapLumaKaris", "mAddNoise", "mReduceBlendNearClamp", "mDynamicAntiGhosting",
       "mDebugCenter", "mDebugToScreenOutput", "mVelBasedAlpha", "mRayTraceAugment"]
INT = ["mColorClampingOrClipping", "mUseVelocityVectors", "mVelocitySampleMode", "mInterpolationMode", "mDebugMode"]
FLT = ["mVarClipGamma", "mUnjitterFactor", "mAlpha", "mMinAlpha", "mMaxAlpha", "mRejectionAlpha", "mNoiseFactor", "mDebugScale",
       "mVelBasedAlphaMax", "mVelBasedAlphaFactor"]
PRIM_BOOL = ["mTaaEnabled", "mSplitScreen", "mResetHistoryOnChange", "mPostProcessEnabled"]
PRIM_INT = ["mSampleDistribution", "mSharpener", "mSplitX", "mFixedJitterIndex", "mJitterSlowMotion"]
PRIM_FLT = ["mSharpenFactor", "mJitterExtraScale", "mJitterRotateDegrees"]


def values():
    v = {"param": [], "primary": {}, "post": {}}
    for i in range(2):
        p = {}
        for k, n in enumerate(B32):
            p[n] = (k + i) % 2
        p["mRayTraceAugment"] = 0
        for k, n in enumerate(INT):
            p[n] = (k + 2 * i) % 3
        for k, n in enumerate(FLT):
            p[n] = round(0.0625 * (k + 1) + 0.3 * i + (1e-7 if k == 2 else 0.0), 8)
        p["mVelBasedAlphaFactor"] = 123456.789 if i else 1e-9
        p["mDebugMask"] = [1.0, 0.0, 0.5 + i, -2.25]
        v["param"].append(p)
    v["primary"] = {"mTaaEnabled": 1, "mSplitScreen": 1, "mResetHistoryOnChange": 0, "mPostProcessEnabled": 1, "mSampleDistribution": 5, "mSharpener": 2,
                    "mSplitX": -17, "mFixedJitterIndex": -1, "mJitterSlowMotion": 4, "mSharpenFactor": 0.35, "mJitterExtraScale": 1.5,
                    "mJitterRotateDegrees": 33.3333, "mDebugSampleOffsets": [[0.25, -0.125], [-0.5, 0.375], [0.0, 1e-3]]}
    v["post"] = {"zoom": 1, "showZoomBox": 0, "zoomSrcLTWH": [960, 540, 10, 10], "zoomDstLTWH": [-3, 0, 200, 200]}
    return v


def cxx_set(v):
    out = []
    for i, p in enumerate(v["param"]):
        for n in B32 + INT:
            out.append(f"s.mParameters[{i}].{n} = {p[n]};")
        for n in FLT:
            out.append(f"s.mParameters[{i}].{n} = {p[n]!r}f;")
        out.append(f"s.mParameters[{i}].mDebugMask = glm::vec4({', '.join(repr(x) + 'f' for x in p['mDebugMask'])});")
    pr = v["primary"]
    for n in PRIM_BOOL:
        out.append(f"s.{n} = {'true' if pr[n] else 'false'};")
    for n in PRIM_INT:
        out.append(f"s.{n} = {pr[n]};")
    for n in PRIM_FLT:
        out.append(f"s.{n} = {pr[n]!r}f;")
    out.append("s.mDebugSampleOffsets = {" + ", ".join(f"glm::vec2({a!r}f, {b!r}f)" for a, b in pr["mDebugSampleOffsets"]) + "};")
    po = v["post"]
    out.append(f"s.mPostProcessPushConstants.zoom = {po['zoom']}; s.mPostProcessPushConstants.showZoomBox = {po['showZoomBox']};")
    out.append(f"s.mPostProcessPushConstants.zoomSrcLTWH = glm::ivec4({', '.join(map(str, po['zoomSrcLTWH']))});")
    out.append(f"s.mPostProcessPushConstants.zoomDstLTWH = glm::ivec4({', '.join(map(str, po['zoomDstLTWH']))});")
    return "\n  ".join(out)


def cxx_dump():
    out = ['printf("{\\"param\\":[");', "for (int i = 0; i < 2; ++i) { auto& p = s.mParameters[i]; printf(\"%s{\", i ? \",\" : \"\");"]
    for n in B32:
        out.append(f'printf("\\"{n}\\":%u,", p.{n});')
    for n in INT:
        out.append(f'printf("\\"{n}\\":%d,", p.{n});')
    for n in FLT:
        out.append(f'printf("\\"{n}\\":%.9g,", p.{n});')
    out.append('printf("\\"mDebugMask\\":[%.9g,%.9g,%.9g,%.9g]}", p.mDebugMask.x, p.mDebugMask.y, p.mDebugMask.z, p.mDebugMask.w); }')
    out.append('printf("],\\"primary\\":{");')
    for n in PRIM_BOOL:
        out.append(f'printf("\\"{n}\\":%d,", s.{n} ? 1 : 0);')
    for n in PRIM_INT:
        out.append(f'printf("\\"{n}\\":%d,", s.{n});')
    for n in PRIM_FLT:
        out.append(f'printf("\\"{n}\\":%.9g,", s.{n});')
    out.append('printf("\\"mDebugSampleOffsets\\":[");')
    out.append('for (size_t i = 0; i < s.mDebugSampleOffsets.size(); ++i) printf("%s[%.9g,%.9g]", i ? "," : "", s.mDebugSampleOffsets[i].x, s.mDebugSampleOffsets[i].y);')
    out.append('auto& pp = s.mPostProcessPushConstants;')
    out.append('printf("]},\\"post\\":{\\"zoom\\":%u,\\"showZoomBox\\":%u,\\"zoomSrcLTWH\\":[%d,%d,%d,%d],\\"zoomDstLTWH\\":[%d,%d,%d,%d]}}\\n", pp.zoom, pp.showZoomBox, '
               'pp.zoomSrcLTWH.x, pp.zoomSrcLTWH.y, pp.zoomSrcLTWH.z, pp.zoomSrcLTWH.w, pp.zoomDstLTWH.x, pp.zoomDstLTWH.y, pp.zoomDstLTWH.z, pp.zoomDstLTWH.w);')
    return "\n  ".join(out)


def extract(src, name):
    i = src.index(f"void {name}(mINI::INIStructure &ini)")
    depth, j = 0, src.index("{", i)
    k = j
    while True:
        depth += {"{": 1, "}": -1}.get(src[k], 0)
        if depth == 0:
            break
        k += 1
    return src[i:k + 1]


def main():
    src = open(os.path.join(REF, "source", "taa.hpp")).read()
    wfn, rfn = extract(src, "writeSettingsToIni"), extract(src, "readSettingsFromIni")
    v = values()
    members = "\n  ".join([f"uint32_t {n} = 0;" for n in B32] + [f"int {n} = 0;" for n in INT] + [f"float {n} = 0;" for n in FLT] + ["glm::vec4 mDebugMask = glm::vec4(0);"])
    prog = f'''
#include <cstdio>
#include <cstdint>
#include <string>
#include <vector>
#include <algorithm>
#include "IniUtil.h"
struct Parameters {{
  {members}
}};
struct PP {{ uint32_t zoom = 0, showZoomBox = 0; glm::ivec4 zoomSrcLTWH = glm::ivec4(0), zoomDstLTWH = glm::ivec4(0); }};
struct Stub {{
  Parameters mParameters[2];
  bool mTaaEnabled = false, mSplitScreen = false, mResetHistoryOnChange = false, mPostProcessEnabled = false;
  int mSampleDistribution = 0, mSharpener = 0, mSplitX = 0, mFixedJitterIndex = 0, mJitterSlowMotion = 0;
  float mSharpenFactor = 0, mJitterExtraScale = 0, mJitterRotateDegrees = 0;
  std::vector<glm::vec2> mDebugSampleOffsets;
  PP mPostProcessPushConstants;
  {wfn}
  {rfn}
}};
int main(int argc, char** argv) {{
  Stub s;
  {cxx_set(v)}
  {{ mINI::INIStructure ini; s.writeSettingsToIni(ini); mINI::INIFile f(argv[1]); f.generate(ini); }}
  {{ mINI::INIStructure ini; mINI::INIFile f(argv[2]); f.read(ini); s.readSettingsFromIni(ini); }}
  {cxx_dump()}
  return 0;
}}
'''
    with tempfile.TemporaryDirectory() as td:
        open(os.path.join(td, "m.cpp"), "w").write(prog)
        exe = os.path.join(td, "m")
        # (the reference compiles IniUtil.cpp behind a precompiled header that brings glm's quaternion type in: force-include it)
        subprocess.check_call(["g++", "-O1", "-std=c++17", "-w", "-include", "glm/gtc/quaternion.hpp", "-I", os.path.join(REF, "source"), "-I", os.path.join(REF, "external", "include"),
                               "-I", os.path.join(REF, "gears_vk", "external", "universal", "include"), os.path.join(td, "m.cpp"),
                               os.path.join(REF, "source", "IniUtil.cpp"), "-o", exe])
        written = os.path.join(HERE, "taa_settings_written.ini")
        out = subprocess.check_output([exe, written, os.path.join(HERE, "taa_settings_input.ini")], text=True)
    json.dump(v, open(os.path.join(HERE, "taa_settings_values.json"), "w"), indent=1)
    json.dump(json.loads(out), open(os.path.join(HERE, "taa_settings_read.json"), "w"), indent=1)
    print("wrote taa_settings_written.ini, taa_settings_values.json, taa_settings_read.json")


if __name__ == "__main__":
    main()
