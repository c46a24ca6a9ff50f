"""Generates tests/golden/matrix_golden.json with the vendored GLM of the reference (run in the build container only): the two matrices
taa<CF>::render() uploads, exactly as source/taa.hpp:993-994 computes them — glm::inverse(P_cur * V_cur) and P_prev * V_prev — for a set of
camera poses. Floats are written as bit patterns. GLM is the reference's arithmetic here (a third-party header it vendors under
gears_vk/external/universal/include/glm, version 0.9.9.9); only the numbers enter the repository."""
import json
import os
import subprocess
import sys
import tempfile

REF = sys.argv[1] if len(sys.argv) > 1 else "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))

PROG = r'''
#include <cmath>
#include <cstdio>
#include <cstring>
#define GLM_ENABLE_EXPERIMENTAL
#include <glm/glm.hpp>
#include <glm/gtc/matrix_transform.hpp>
static unsigned bits(float f) { unsigned u; memcpy(&u, &f, 4); return u; }
static void dump(const char* k, const glm::mat4& m, bool comma) {
  printf("\"%s\":[", k);
  for (int c = 0; c < 4; ++c) for (int r = 0; r < 4; ++r) printf("%s%u", (c || r) ? "," : "", bits(m[c][r]));
  printf("]%s", comma ? "," : "");
}
static glm::mat4 persp(float aspect) {  // gvk's perspective (camera.cpp:182-191): fov 60 deg, near 0.1, far 100, y down, depth 0..1
  const float f = 1.0f / std::tan(0.5f * 1.04719755f), n = 0.1f, fr = 100.f;
  glm::mat4 P(0.f); P[0][0] = f / aspect; P[1][1] = -f; P[2][2] = fr / (fr - n); P[2][3] = 1.f; P[3][2] = -(fr * n) / (fr - n);
  return P;
}
int main() {
  printf("[");
  for (int i = 0; i < 12; ++i) {
    const float a = 0.37f * (float)i, b = 0.11f * (float)i;
    const glm::vec3 eye(3.0f * std::cos(a), 1.0f + 0.3f * (float)i, 3.0f * std::sin(a));
    const glm::vec3 eye2 = eye + glm::vec3(0.013f * (float)(i + 1), -0.004f, 0.021f);
    const glm::mat4 Vc = glm::lookAt(eye, glm::vec3(0.1f * (float)i, 0.5f, -0.2f), glm::vec3(std::sin(b) * 0.1f, 1.0f, 0.0f));
    const glm::mat4 Vp = glm::lookAt(eye2, glm::vec3(0.1f * (float)i + 0.01f, 0.5f, -0.2f), glm::vec3(0.0f, 1.0f, 0.0f));
    const glm::mat4 Pc = persp(i % 2 ? 16.0f / 9.0f : 4.0f / 3.0f), Pp = persp(i % 3 ? 16.0f / 9.0f : 1.0f);
    const glm::mat4 inv = glm::inverse(Pc * Vc);     // taa.hpp:993
    const glm::mat4 hist = Pp * Vp;                  // taa.hpp:994
    printf("%s{", i ? "," : "");
    dump("proj_cur", Pc, true); dump("view_cur", Vc, true); dump("proj_prev", Pp, true); dump("view_prev", Vp, true);
    dump("inverse_view_proj", inv, true); dump("history_view_proj", hist, false);
    printf("}");
  }
  printf("]\n");
  return 0;
}
'''

with tempfile.TemporaryDirectory() as td:
    open(os.path.join(td, "m.cpp"), "w").write(PROG)
    exe = os.path.join(td, "m")
    subprocess.check_call(["g++", "-O1", "-std=c++17", "-w", "-ffp-contract=off", "-I", os.path.join(REF, "gears_vk", "external", "universal", "include"),
                           os.path.join(td, "m.cpp"), "-o", exe])
    out = json.loads(subprocess.check_output([exe], text=True))
json.dump(out, open(os.path.join(HERE, "matrix_golden.json"), "w"))
print("wrote matrix_golden.json:", len(out), "camera poses")
