// Instruction-throughput microbenchmarks that drive the kernel design (run on the B200 box).
#include <cstdio>
#include <cuda_runtime.h>
#include <stdint.h>
#define ITERS 4096
template <int OP> __global__ void __launch_bounds__(1024) k(double* out, int n)
{
	__shared__ unsigned int sh[4096];
	__shared__ double shd[2048];
	const int tid = threadIdx.x;
	for (int i = tid; i < 4096; i += blockDim.x) sh[i] = i;
	for (int i = tid; i < 2048; i += blockDim.x) shd[i] = i;
	__syncthreads();
	double a0 = tid * 1e-3, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4 = a0 + 4, a5 = a0 + 5, a6 = a0 + 6, a7 = a0 + 7;
	float f0 = tid * 1e-3f, f1 = f0 + 1, f2 = f0 + 2, f3 = f0 + 3, f4 = f0 + 4, f5 = f0 + 5, f6 = f0 + 6, f7 = f0 + 7;
	unsigned u0 = tid, u1 = tid + 1, u2 = tid + 2, u3 = tid + 3;
	const double m = 1.0000001, c = 1e-9;
	const float mf = 1.0000001f, cf = 1e-9f;
	for (int it = 0; it < n; ++it) {
		if (OP == 0) { a0 = fma(a0, m, c); a1 = fma(a1, m, c); a2 = fma(a2, m, c); a3 = fma(a3, m, c); a4 = fma(a4, m, c); a5 = fma(a5, m, c); a6 = fma(a6, m, c); a7 = fma(a7, m, c); }
		if (OP == 1) { f0 = fmaf(f0, mf, cf); f1 = fmaf(f1, mf, cf); f2 = fmaf(f2, mf, cf); f3 = fmaf(f3, mf, cf); f4 = fmaf(f4, mf, cf); f5 = fmaf(f5, mf, cf); f6 = fmaf(f6, mf, cf); f7 = fmaf(f7, mf, cf); }
		if (OP == 2) { a0 += (double)f0; a1 += (double)f1; a2 += (double)f2; a3 += (double)f3; f0 += cf; f1 += cf; f2 += cf; f3 += cf; }  // F2F + DADD + FADD
		if (OP == 3) { a0 += c; a1 += c; a2 += c; a3 += c; a4 += c; a5 += c; a6 += c; a7 += c; }  // DADD
		if (OP == 4) { atomicAdd(&sh[(u0 * 33) & 4095], 1u); u0 += 7; atomicAdd(&sh[(u1 * 33) & 4095], 1u); u1 += 7; }  // spread smem atomics (2 per iter)
		if (OP == 5) { atomicAdd(&sh[(it & 7)], 1u); atomicAdd(&sh[8 + (it & 7)], 1u); }  // same address within warp
		if (OP == 6) { u0 += __shfl_xor_sync(0xffffffffu, u0, 1); u1 += __shfl_xor_sync(0xffffffffu, u1, 2); u2 += __shfl_xor_sync(0xffffffffu, u2, 4); u3 += __shfl_xor_sync(0xffffffffu, u3, 8); }
		if (OP == 7) { u0 += sh[(u0 + tid) & 4095]; u1 += sh[(u1 + tid) & 4095]; u2 += sh[(u2 + tid) & 4095]; u3 += sh[(u3 + tid) & 4095]; }  // dependent LDS (conflict-free)
		if (OP == 8) { u0 += __match_any_sync(0xffffffffu, u0 >> 3); u1 += __match_any_sync(0xffffffffu, u1 >> 3); }
		if (OP == 9) { a0 = fmin(a0, a1 + c); a1 = fmax(a1, a2); a2 = fmin(a2, a3); a3 = fmax(a3, a0); }  // DMNMX-ish
		if (OP == 10) { u0 += (a0 < a1) ? 1 : 0; u1 += (a1 < a2) ? 1 : 0; u2 += (a2 < a3) ? 1 : 0; u3 += (a3 < a0) ? 1 : 0; a0 += c; }  // DSETP
		if (OP == 11) { u0 = u0 * 3 + 1; u1 = u1 * 3 + 1; u2 = u2 * 3 + 1; u3 = u3 * 3 + 1; u0 ^= u1; u2 ^= u3; }  // int ALU
		if (OP == 12) { u0 += __popc(__ballot_sync(0xffffffffu, u0 & 1)); u1 += __popc(__ballot_sync(0xffffffffu, u1 & 2)); }
		if (OP == 13) { atomicAdd(&shd[(u0 * 33) & 2047], 1.0); u0 += 7; }  // f64 smem atomic (CAS loop)
		if (OP == 14) { f0 += (float)(int)(f1 * mf); f1 += cf; f2 += (float)(int)(f3 * mf); f3 += cf; }  // F2I + I2F
		if (OP == 15) { uint4 v = *reinterpret_cast<uint4*>(&sh[((u0 + tid) * 4) & 4095]); u0 += v.x + v.y + v.z + v.w; }  // LDS.128
	}
	out[blockIdx.x * blockDim.x + tid] = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7 + f0 + f1 + f2 + f3 + f4 + f5 + f6 + f7 + u0 + u1 + u2 + u3 + sh[tid] + shd[tid];
}
template <int OP> void run(const char* name, double ops_per_iter)
{
	double* out; cudaMalloc(&out, 148 * 2 * 1024 * sizeof(double));
	cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
	k<OP><<<148 * 2, 1024>>>(out, 64);
	cudaEventRecord(e0);
	k<OP><<<148 * 2, 1024>>>(out, ITERS);
	cudaEventRecord(e1); cudaEventSynchronize(e1);
	float ms; cudaEventElapsedTime(&ms, e0, e1);
	double lane_ops = 148.0 * 2 * 1024 * ITERS * ops_per_iter;
	double per_clk_sm = lane_ops / (ms * 1e-3) / 148 / 1.965e9;
	printf("%-28s %8.3f ms  %8.1f lane-ops/clk/SM  (%.2f warp-instr/clk/SM)\n", name, ms, per_clk_sm, per_clk_sm / 32);
	cudaFree(out);
}
int main()
{
	run<0>("DFMA", 8); run<1>("FFMA", 8); run<2>("F2F.F64.F32+DADD+FADD", 4); run<3>("DADD", 8);
	run<4>("ATOMS spread", 2); run<5>("ATOMS same-addr", 2); run<6>("SHFL", 4); run<7>("LDS dep", 4);
	run<8>("MATCH.ANY", 2); run<9>("DMNMX", 4); run<10>("DSETP", 4); run<11>("IMAD/LOP", 6);
	run<12>("BALLOT+POPC", 2); run<13>("ATOMS f64 spread", 1); run<14>("F2I+I2F", 2); run<15>("LDS.128", 1);
	return 0;
}
