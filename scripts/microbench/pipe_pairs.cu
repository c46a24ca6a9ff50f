// Which instruction kinds share an issue pipe on B200? Times loops of A-only, B-only and A+B interleaved (CUDA events, whole grid).
// Output unit: warp-instructions per SM clock per SMSP, assuming the SM clock given on the command line (MHz, default 1965).
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
typedef unsigned long long u64;
#define ITER 1024

enum { FADD, FMUL, FFMA, FADD2, FMUL2, FFMA2, CVT_H2F, F2FP, FMNMX, IMAD, IADD3, LOP3, FSETP_SEL, NONE };
static const char* NAMES[] = {"FADD", "FMUL", "FFMA", "FADD2", "FMUL2", "FFMA2", "cvt.f32.f16", "cvt.f16x2.f32", "FMNMX", "IMAD", "IADD3", "LOP3", "FSETP+FSEL", "-"};

template <int K>
__device__ __forceinline__ void op(float& a, u64& p, unsigned& h, int& q, float b, float c, u64 pb, int qi) {
	if (K == FADD) asm volatile("add.rn.f32 %0, %0, %1;" : "+f"(a) : "f"(b));
	if (K == FMUL) asm volatile("mul.rn.f32 %0, %0, %1;" : "+f"(a) : "f"(c));
	if (K == FFMA) asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(a) : "f"(c), "f"(b));
	if (K == FADD2) asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(p) : "l"(pb));
	if (K == FMUL2) asm volatile("mul.rn.f32x2 %0, %0, %1;" : "+l"(p) : "l"(pb));
	if (K == FFMA2) asm volatile("fma.rn.f32x2 %0, %0, %1, %1;" : "+l"(p) : "l"(pb));
	if (K == CVT_H2F) { float f; asm volatile("{.reg .b16 lo, hi; mov.b32 {lo, hi}, %1; cvt.f32.f16 %0, lo;}" : "=f"(f) : "r"(h)); h = __float_as_uint(f); }
	if (K == F2FP) { asm volatile("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(h) : "f"(__uint_as_float(h)), "f"(b)); }
	if (K == FMNMX) asm volatile("min.f32 %0, %0, %1;" : "+f"(a) : "f"(b));
	if (K == IMAD) asm volatile("mad.lo.s32 %0, %0, %1, %1;" : "+r"(q) : "r"(qi));
	if (K == IADD3) asm volatile("add.s32 %0, %0, %1;" : "+r"(q) : "r"(qi));
	if (K == LOP3) asm volatile("xor.b32 %0, %0, %1;" : "+r"(q) : "r"(qi));
	if (K == FSETP_SEL) asm volatile("{.reg .pred pp; setp.gt.f32 pp, %0, %1; selp.f32 %0, %1, %0, pp;}" : "+f"(a) : "f"(b));
}

template <int A, int B>
__global__ void __launch_bounds__(256) bench(float* out, float seed, int qi) {
	float a[8], a2[8];
	u64 p[8], p2[8];
	unsigned h[8], h2[8];
	int q[8], q2[8];
	for (int i = 0; i < 8; ++i) {
		a[i] = seed + i + threadIdx.x; a2[i] = a[i] * 3.f;
		float2 t = make_float2(a[i], a[i] * 0.5f); p[i] = *reinterpret_cast<u64*>(&t); p2[i] = p[i] + 7;
		h[i] = __float_as_uint(a[i]); h2[i] = h[i] ^ 0x55;
		q[i] = threadIdx.x + i; q2[i] = q[i] * 3;
	}
	float b = seed * 1.0001f, c = seed * 0.99f;
	float2 bb = make_float2(b, c);
	u64 pb = *reinterpret_cast<u64*>(&bb);
#pragma unroll 1
	for (int it = 0; it < ITER; ++it) {
#pragma unroll
		for (int r = 0; r < 4; ++r) {
#pragma unroll
			for (int i = 0; i < 8; ++i) {
				op<A>(a[i], p[i], h[i], q[i], b, c, pb, qi);
				if (B != NONE) op<B>(a2[i], p2[i], h2[i], q2[i], b, c, pb, qi);
			}
		}
	}
	float s = 0;
	for (int i = 0; i < 8; ++i) {
		float2 t = *reinterpret_cast<float2*>(&p[i]), t2 = *reinterpret_cast<float2*>(&p2[i]);
		s += a[i] + a2[i] + t.x + t.y + t2.x + t2.y + __uint_as_float(h[i]) + __uint_as_float(h2[i]) + q[i] + q2[i];
	}
	out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

static double g_mhz = 1965.0;
template <int A, int B>
void run() {
	const int sms = 148, bps = 4;
	float* out;
	cudaMalloc(&out, sms * bps * 256 * 4);
	cudaEvent_t e0, e1;
	cudaEventCreate(&e0); cudaEventCreate(&e1);
	bench<A, B><<<sms * bps, 256>>>(out, 1.0f, 3);
	cudaDeviceSynchronize();
	cudaEventRecord(e0);
	bench<A, B><<<sms * bps, 256>>>(out, 1.0f, 3);
	cudaEventRecord(e1);
	cudaDeviceSynchronize();
	float ms;
	cudaEventElapsedTime(&ms, e0, e1);
	double clk = ms * 1e-3 * g_mhz * 1e6;
	double n = (B == NONE ? 1.0 : 2.0) * 8 /*warps per SMSP*/ * ITER * 32.0;
	printf("%-14s + %-14s : %5.3f warp-inst/clk/SMSP  (%.3f ms)\n", NAMES[A], NAMES[B], n / clk, ms);
	cudaFree(out);
}

int main(int argc, char** argv) {
	if (argc > 1) g_mhz = atof(argv[1]);
	run<FADD, NONE>(); run<FMUL, NONE>(); run<FFMA, NONE>(); run<FADD2, NONE>(); run<FMUL2, NONE>(); run<FFMA2, NONE>();
	run<CVT_H2F, NONE>(); run<F2FP, NONE>(); run<FMNMX, NONE>(); run<IMAD, NONE>(); run<IADD3, NONE>(); run<LOP3, NONE>(); run<FSETP_SEL, NONE>();
	run<FADD, FMUL>(); run<FADD, FFMA>(); run<FMUL, FFMA>(); run<FADD, FMNMX>(); run<FMUL, FMNMX>(); run<FFMA, FMNMX>();
	run<FADD, IMAD>(); run<FADD, IADD3>(); run<FMUL, IADD3>(); run<FFMA, IADD3>(); run<FFMA, IMAD>();
	run<FADD2, FADD>(); run<FADD2, FMUL>(); run<FADD2, FFMA>(); run<FADD2, FMNMX>(); run<FADD2, IADD3>(); run<FADD2, IMAD>();
	run<FMUL2, FADD>(); run<FMUL2, FMUL>(); run<FMUL2, FMNMX>(); run<FMUL2, IADD3>();
	run<FFMA2, FADD>(); run<FFMA2, FMUL>(); run<FFMA2, FFMA>(); run<FFMA2, FMNMX>(); run<FFMA2, IADD3>(); run<FFMA2, IMAD>();
	run<FADD2, FMUL2>(); run<FADD2, FFMA2>(); run<FMUL2, FFMA2>();
	run<CVT_H2F, FADD>(); run<CVT_H2F, FMUL>(); run<CVT_H2F, FFMA2>(); run<CVT_H2F, FADD2>(); run<CVT_H2F, IADD3>();
	run<F2FP, FADD>(); run<F2FP, FMUL>();
	return 0;
}
