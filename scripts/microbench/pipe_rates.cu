// Issue-rate microbenchmark for the instruction kinds the resolve kernels are made of (B200, sm_100a).
// Prints warp-instructions per clock per SM for each kind. Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 pipe_rates.cu
#include <cstdio>
#include <cuda_fp16.h>
#include <cuda_runtime.h>

#define ITER 2048
typedef unsigned long long u64;

template <int KIND>
__global__ void __launch_bounds__(256) bench(float* out, float seed, long long* cycles) {
	float a[8];
	u64 p[8];
	for (int i = 0; i < 8; ++i) { a[i] = seed + i + threadIdx.x; float2 t = make_float2(a[i], a[i] * 0.5f); p[i] = *reinterpret_cast<u64*>(&t); }
	float b = seed * 1.0001f, c = seed * 0.5f;
	float2 bb = make_float2(b, c);
	u64 pb = *reinterpret_cast<u64*>(&bb);
	unsigned int h[8];
	for (int i = 0; i < 8; ++i) h[i] = __float_as_uint(a[i]);
	long long t0 = clock64();
#pragma unroll 1
	for (int it = 0; it < ITER; ++it) {
#pragma unroll
		for (int i = 0; i < 8; ++i) {
			if (KIND == 0) asm volatile("add.rn.f32 %0, %0, %1;" : "+f"(a[i]) : "f"(b));
			if (KIND == 1) asm volatile("mul.rn.f32 %0, %0, %1;" : "+f"(a[i]) : "f"(b));
			if (KIND == 2) asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(a[i]) : "f"(b), "f"(c));
			if (KIND == 3) asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(p[i]) : "l"(pb));
			if (KIND == 4) asm volatile("mul.rn.f32x2 %0, %0, %1;" : "+l"(p[i]) : "l"(pb));
			if (KIND == 5) asm volatile("fma.rn.f32x2 %0, %0, %1, %1;" : "+l"(p[i]) : "l"(pb));
			if (KIND == 6) { float f; asm volatile("{.reg .b16 lo, hi; mov.b32 {lo, hi}, %1; cvt.f32.f16 %0, lo;}" : "=f"(f) : "r"(h[i])); h[i] = __float_as_uint(f); }
			if (KIND == 7) { asm volatile("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(h[i]) : "f"(a[i]), "f"(b)); a[i] = __uint_as_float(h[i]); }
			if (KIND == 8) asm volatile("min.f32 %0, %0, %1;" : "+f"(a[i]) : "f"(b));
			if (KIND == 9) { asm volatile("add.rn.f32 %0, %0, %1;" : "+f"(a[i]) : "f"(b)); asm volatile("fma.rn.f32x2 %0, %0, %1, %1;" : "+l"(p[i]) : "l"(pb)); }
			if (KIND == 10) { asm volatile("add.rn.f32 %0, %0, %1;" : "+f"(a[i]) : "f"(b)); asm volatile("mul.rn.f32 %0, %0, %1;" : "+f"(a[(i + 4) & 7]) : "f"(c)); }
			if (KIND == 11) asm volatile("rcp.approx.f32 %0, %0;" : "+f"(a[i]));
			if (KIND == 12) { int q; asm volatile("cvt.rzi.s32.f32 %0, %1;" : "=r"(q) : "f"(a[i])); a[i] = __int_as_float(q); }
			if (KIND == 13) asm volatile("cvt.rmi.f32.f32 %0, %0;" : "+f"(a[i]));
			if (KIND == 14) { asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(p[i]) : "l"(pb)); asm volatile("min.f32 %0, %0, %1;" : "+f"(a[i]) : "f"(b)); }
			if (KIND == 15) { asm volatile("mul.rn.f32x2 %0, %0, %1;" : "+l"(p[i]) : "l"(pb)); asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(p[(i + 4) & 7]) : "l"(pb)); }
			if (KIND == 16) { float f; asm volatile("{.reg .b16 lo, hi; mov.b32 {lo, hi}, %1; cvt.f32.f16 %0, lo;}" : "=f"(f) : "r"(h[i])); h[i] = __float_as_uint(f); asm volatile("fma.rn.f32x2 %0, %0, %1, %1;" : "+l"(p[i]) : "l"(pb)); }
		}
	}
	long long t1 = clock64();
	float s = 0;
	for (int i = 0; i < 8; ++i) { float2 t = *reinterpret_cast<float2*>(&p[i]); s += a[i] + t.x + t.y + __uint_as_float(h[i]); }
	out[blockIdx.x * blockDim.x + threadIdx.x] = s;
	if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

template <int KIND>
void run(const char* name, int per_iter) {
	int sms = 148, bps = 4;  // 4 blocks x 8 warps = 32 warps / SM
	float* out;
	long long* cyc;
	cudaMalloc(&out, sms * bps * 256 * 4);
	cudaMalloc(&cyc, sms * bps * 8);
	bench<KIND><<<sms * bps, 256>>>(out, 1.0f, cyc);
	cudaDeviceSynchronize();
	bench<KIND><<<sms * bps, 256>>>(out, 1.0f, cyc);
	cudaDeviceSynchronize();
	long long h[148 * 4];
	cudaMemcpy(h, cyc, sizeof h, cudaMemcpyDeviceToHost);
	double avg = 0;
	for (auto v : h) avg += v;
	avg /= (sms * bps);
	double warp_inst_per_sm = (double)bps * 8 * ITER * 8 * per_iter;
	printf("%-34s %6.3f warp-inst/clk/SM  (%.0f clk)\n", name, warp_inst_per_sm / avg, avg);
	cudaFree(out);
	cudaFree(cyc);
}

int main() {
	run<0>("FADD", 1);
	run<1>("FMUL", 1);
	run<2>("FFMA", 1);
	run<3>("FADD2 (f32x2)", 1);
	run<4>("FMUL2 (f32x2)", 1);
	run<5>("FFMA2 (f32x2)", 1);
	run<6>("cvt.f32.f16 (HADD2.F32)", 1);
	run<7>("cvt.rn.f16x2.f32 (F2FP)", 1);
	run<8>("FMNMX", 1);
	run<9>("FADD + FFMA2", 2);
	run<10>("FADD + FMUL", 2);
	run<11>("MUFU.RCP", 1);
	run<12>("F2I", 1);
	run<13>("FRND.FLOOR", 1);
	run<14>("FADD2 + FMNMX", 2);
	run<15>("FMUL2 + FADD2", 2);
	run<16>("cvt.f32.f16 + FFMA2", 2);
	return 0;
}
