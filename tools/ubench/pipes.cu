// pipes.cu - issue/pipe throughput of the instruction mixes k_fused is made of (B200, sm_100a).
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o pipes pipes.cu && ./pipes
// Each test: 148 CTAs x 640 threads (20 warps/SM like k_fused), ITER x UNROLL independent chains per thread.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

#define ITER 2048

template <int MODE>
__global__ void __launch_bounds__(640, 1) k(float *out, uint32_t seed)
{
	// 8 independent chains per type
	float f[8];
	float2 g[8];
	uint32_t u[8];
#pragma unroll
	for (int i = 0; i < 8; i++) {
		f[i] = (float)(threadIdx.x + i) * 1e-3f;
		g[i] = make_float2(f[i], f[i] + 1.0f);
		u[i] = threadIdx.x * 2654435761u + i + seed;
	}
	const float c1 = 0.999f + seed * 1e-9f, c2 = 1e-3f;
	const float2 c12 = make_float2(c1, c1), c22 = make_float2(c2, c2);
	for (int it = 0; it < ITER; it++) {
#pragma unroll
		for (int i = 0; i < 8; i++) {
			if (MODE == 0) f[i] = __fmaf_rn(f[i], c1, c2);                       // FFMA
			if (MODE == 1) g[i] = __ffma2_rn(g[i], c12, c22);                    // FFMA2
			if (MODE == 2) g[i] = __fadd2_rn(g[i], c22);                         // FADD2
			if (MODE == 3) u[i] = __byte_perm(u[i], seed, 0x7610 + (i & 1));     // PRMT
			if (MODE == 4) u[i] = __funnelshift_l(u[(i + 1) & 7], u[i], 1);      // SHF
			if (MODE == 5) u[i] = (u[i] & seed) ^ u[(i + 1) & 7];                // LOP3
			if (MODE == 6) { g[i] = __ffma2_rn(g[i], c12, c22); u[i] = __funnelshift_l(u[(i + 1) & 7], u[i], 1); }  // FFMA2 + SHF
			if (MODE == 7) { f[i] = __fmaf_rn(f[i], c1, c2); u[i] = __funnelshift_l(u[(i + 1) & 7], u[i], 1); }     // FFMA + SHF
			if (MODE == 8) { g[i] = __ffma2_rn(g[i], c12, c22); f[i] = __fmaf_rn(f[i], c1, c2); }                   // FFMA2 + FFMA
			if (MODE == 9) { g[i] = __ffma2_rn(g[i], c12, c22); f[i] = __fmaf_rn(f[i], c1, c2);
			                 u[i] = __funnelshift_l(u[(i + 1) & 7], u[i], 1); }                                        // FFMA2 + FFMA + SHF
			if (MODE == 10) { f[i] = __fsub_rn(f[i], g[i].x); u[i] = __funnelshift_l(__float_as_uint(f[i]), u[i], 1); } // FADD -> SHF dependent (bits stage)
			if (MODE == 11) u[i] = u[i] + u[(i + 1) & 7] + seed;                  // IADD3
			if (MODE == 12) { g[i] = __ffma2_rn(g[i], c12, c22); u[i] = __byte_perm(u[i], seed, 0x7610 + (i & 1)); }   // FFMA2 + PRMT
			if (MODE == 13) { g[i] = __ffma2_rn(g[i], c12, c22); g[(i + 4) & 7] = __fadd2_rn(g[(i + 4) & 7], c22); }    // FFMA2 + FADD2
			if (MODE == 14) f[i] = __fmaf_rn(f[i], f[(i + 1) & 7], f[(i + 2) & 7]); // FFMA 3 distinct regs
			if (MODE == 15) g[i] = __ffma2_rn(g[i], g[(i + 1) & 7], g[(i + 2) & 7]); // FFMA2 3 distinct reg pairs
		}
		if (MODE == 16) {  // the same 8 FFMA2 + 8 FFMA as mode 8, but in two blocks instead of alternating
#pragma unroll
			for (int i = 0; i < 8; i++) g[i] = __ffma2_rn(g[i], c12, c22);
#pragma unroll
			for (int i = 0; i < 8; i++) f[i] = __fmaf_rn(f[i], c1, c2);
		}
		if (MODE == 17) {  // 8 FFMA2 + 8 FFMA in groups of two
#pragma unroll
			for (int i = 0; i < 8; i += 2) {
				g[i] = __ffma2_rn(g[i], c12, c22); g[i + 1] = __ffma2_rn(g[i + 1], c12, c22);
				f[i] = __fmaf_rn(f[i], c1, c2); f[i + 1] = __fmaf_rn(f[i + 1], c1, c2);
			}
		}
		if (MODE == 18) {  // squares as k_fused has them: FMUL + FFMA (scalar) per FFMA2 x 7
#pragma unroll
			for (int i = 0; i < 8; i++) {
				g[i] = __ffma2_rn(g[i], c12, c22);
				if (i == 7) { f[0] = __fmul_rn(f[0], c1); f[1] = __fmaf_rn(f[1], c1, c2); }
			}
		}
		if (MODE == 19) {  // the same with one packed pair instead of the two scalars
#pragma unroll
			for (int i = 0; i < 8; i++)
				g[i] = __ffma2_rn(g[i], c12, c22);
		}
	}
	float acc = 0;
#pragma unroll
	for (int i = 0; i < 8; i++)
		acc += f[i] + g[i].x + g[i].y + __uint_as_float(u[i] & 0x3FFFFFFF);
	if (acc == 12345.678f)
		out[0] = acc;
}

template <int MODE>
void run(const char *name, int instr_per_iter, float *d)
{
	cudaEvent_t a, b;
	cudaEventCreate(&a);
	cudaEventCreate(&b);
	k<MODE><<<148, 640>>>(d, 1);
	cudaDeviceSynchronize();
	cudaEventRecord(a);
	k<MODE><<<148, 640>>>(d, 2);
	cudaEventRecord(b);
	cudaDeviceSynchronize();
	float ms;
	cudaEventElapsedTime(&ms, a, b);
	int clk_khz;
	cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, 0);
	const double warp_instr = 20.0 * ITER * 8 * instr_per_iter;          // per SM
	const double cycles = ms * 1e-3 * clk_khz * 1e3;                      // at the nominal max clock
	printf("%-28s %8.3f ms  %6.3f warp-instr/clk/SM (of 4 issue slots)  [%d instr per step]\n", name, ms, warp_instr / cycles,
	       instr_per_iter);
}

int main()
{
	float *d;
	cudaMalloc(&d, 4);
	run<0>("FFMA (reg,const,const)", 1, d);
	run<14>("FFMA (3 regs)", 1, d);
	run<1>("FFMA2 (pair,const,const)", 1, d);
	run<15>("FFMA2 (3 reg pairs)", 1, d);
	run<2>("FADD2", 1, d);
	run<3>("PRMT", 1, d);
	run<4>("SHF", 1, d);
	run<5>("LOP3", 1, d);
	run<11>("IADD3", 1, d);
	run<6>("FFMA2 + SHF", 2, d);
	run<7>("FFMA + SHF", 2, d);
	run<8>("FFMA2 + FFMA", 2, d);
	run<9>("FFMA2 + FFMA + SHF", 3, d);
	run<10>("FADD -> SHF (dependent)", 2, d);
	run<12>("FFMA2 + PRMT", 2, d);
	run<13>("FFMA2 + FADD2", 2, d);
	run<16>("8 FFMA2 then 8 FFMA", 2, d);
	run<17>("FFMA2 x2, FFMA x2", 2, d);
	run<18>("7 FFMA2 + FMUL + FFMA", 1, d);
	run<19>("8 FFMA2", 1, d);
	return 0;
}
