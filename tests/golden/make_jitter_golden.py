"""Generates tests/golden/jitter_golden.json from REFERENCE code compiled as-is (run in the build container only).

The bodies of taa<CF>::get_jitter_offset_for_frame and get_jittered_projection_matrix (source/taa.hpp:150-253) and of helpers::halton /
halton_2_3 (source/helper_functions.hpp:9-26) are extracted at run time and compiled inside a stub class that has the members they touch,
against the GLM the reference vendors. The function keeps its patterns in function-local statics sized by the input resolution, so the
program is run once per resolution. Floats are written as their bit patterns. Nothing of the reference is copied into the repository:
only the numbers are.
"""
import json
import os
import re
import subprocess
import sys
import tempfile

REF = sys.argv[1] if len(sys.argv) > 1 else "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))


def block(src, start_pat):
    i = re.search(start_pat, src).start()
    j = src.index("{", i)
    depth, k = 0, j
    while True:
        depth += {"{": 1, "}": -1}.get(src[k], 0)
        if depth == 0:
            return src[i:k + 1]
        k += 1


def main():
    taa = open(os.path.join(REF, "source", "taa.hpp")).read()
    hlp = open(os.path.join(REF, "source", "helper_functions.hpp")).read()
    halton = block(hlp, r"static float halton\(int i, int b\)")
    halton23 = "template <size_t Len>\n" + block(hlp, r"static std::array<glm::vec2, Len> halton_2_3\(glm::vec2 aScale\)")
    jit = block(taa, r"glm::vec2 get_jitter_offset_for_frame\(")
    proj = block(taa, r"glm::mat4 get_jittered_projection_matrix\(glm::mat4 aProjMatrix, glm::vec2 &out_xyOffset")
    prog = r'''
#include <array>
#include <cassert>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#define GLM_ENABLE_EXPERIMENTAL
#include <glm/glm.hpp>
#include <glm/gtx/transform.hpp>
namespace gvk { struct window { using frame_id_t = int64_t; }; }
namespace avk { template<typename Type, typename ... T> constexpr auto make_array(T&&... t)->std::array<Type, sizeof...(T)> { return { {std::forward<T>(t)...} }; } }
namespace helpers {
''' + halton + "\n" + halton23 + r'''
}
struct Stub {
  glm::uvec2 mInputResolution;
  bool mTaaEnabled = true;
  int mSampleDistribution = 1, mFixedJitterIndex = -1, mJitterSlowMotion = 1;
  float mJitterExtraScale = 1.f, mJitterRotateDegrees = 0.f;
  std::vector<glm::vec2> mDebugSampleOffsets;
''' + jit + "\n" + proj + r'''
};
static unsigned bits(float f) { unsigned u; memcpy(&u, &f, 4); return u; }
int main(int argc, char** argv) {
  Stub s;
  s.mInputResolution = glm::uvec2(atoi(argv[1]), atoi(argv[2]));
  s.mDebugSampleOffsets = { glm::vec2(0.25f, -0.125f), glm::vec2(-0.5f, 0.375f), glm::vec2(0.0f, 0.001f) };
  struct Set { int fixed, slow; float scale, rot; };
  const Set sets[] = { {-1, 1, 1.f, 0.f}, {-1, 3, 1.5f, 0.f}, {2, 1, 1.f, 30.f}, {-1, 2, 0.75f, -77.5f} };
  printf("{\"w\":%s,\"h\":%s,\"cases\":[", argv[1], argv[2]);
  bool first = true;
  for (int dist = 0; dist < 6; ++dist) for (const Set& t : sets) {
    s.mSampleDistribution = dist; s.mFixedJitterIndex = t.fixed; s.mJitterSlowMotion = t.slow; s.mJitterExtraScale = t.scale; s.mJitterRotateDegrees = t.rot;
    printf("%s{\"dist\":%d,\"fixed\":%d,\"slow\":%d,\"scale\":%u,\"rot\":%u,\"n\":", first ? "" : ",", dist, t.fixed, t.slow, bits(t.scale), bits(t.rot));
    first = false;
    size_t n = 0; s.get_jitter_offset_for_frame(0, nullptr, &n);
    printf("%zu,\"xy\":[", n);
    for (int f = 0; f < 20; ++f) { glm::vec2 p = s.get_jitter_offset_for_frame(f); printf("%s%u,%u", f ? "," : "", bits(p.x), bits(p.y)); }
    printf("]}");
  }
  // get_jittered_projection_matrix: translate(jitter) * P for gvk's perspective (fov 60 deg, near 0.1, far 100) and an arbitrary matrix
  printf("],\"proj\":[");
  s.mSampleDistribution = 2; s.mFixedJitterIndex = -1; s.mJitterSlowMotion = 1; s.mJitterExtraScale = 1.f; s.mJitterRotateDegrees = 0.f;
  const float aspect = (float)s.mInputResolution.x / (float)s.mInputResolution.y, f = 1.0f / std::tan(0.5f * 1.04719755f), n = 0.1f, fr = 100.f;
  glm::mat4 P(0.f); P[0][0] = f / aspect; P[1][1] = -f; P[2][2] = fr / (n - fr) * -1.f; P[2][3] = 1.f; P[3][2] = -(fr * n) / (fr - n);
  glm::mat4 Q(0.f); for (int c = 0; c < 4; ++c) for (int r = 0; r < 4; ++r) Q[c][r] = 0.125f * (float)(c * 4 + r) - 0.7f + (c == r ? 1.0f : 0.0f);
  const glm::mat4 mats[2] = { P, Q };
  for (int m = 0; m < 2; ++m) for (int fidx = 0; fidx < 4; ++fidx) {
    glm::vec2 off; glm::mat4 J = s.get_jittered_projection_matrix(mats[m], off, fidx + 3);
    printf("%s{\"frame\":%d,\"in\":[", (m || fidx) ? "," : "", fidx + 3);
    for (int c = 0; c < 4; ++c) for (int r = 0; r < 4; ++r) printf("%s%u", (c || r) ? "," : "", bits(mats[m][c][r]));
    printf("],\"off\":[%u,%u],\"out\":[", bits(off.x), bits(off.y));
    for (int c = 0; c < 4; ++c) for (int r = 0; r < 4; ++r) printf("%s%u", (c || r) ? "," : "", bits(J[c][r]));
    printf("]}");
  }
  printf("]}\n");
  return 0;
}
'''
    out = []
    with tempfile.TemporaryDirectory() as td:
        open(os.path.join(td, "j.cpp"), "w").write(prog)
        exe = os.path.join(td, "j")
        subprocess.check_call(["g++", "-O1", "-std=c++17", "-w", "-ffp-contract=off", "-I", os.path.join(REF, "gears_vk", "external", "universal", "include"),
                               os.path.join(td, "j.cpp"), "-o", exe])
        for w, h in ((1920, 1080), (3840, 2160), (320, 180), (1000, 333)):
            out.append(json.loads(subprocess.check_output([exe, str(w), str(h)], text=True)))
    json.dump({"debug_offsets": [[0.25, -0.125], [-0.5, 0.375], [0.0, 0.001]], "resolutions": out}, open(os.path.join(HERE, "jitter_golden.json"), "w"))
    print("wrote jitter_golden.json:", sum(len(r["cases"]) for r in out), "pattern cases,", sum(len(r["proj"]) for r in out), "matrices")


if __name__ == "__main__":
    main()
