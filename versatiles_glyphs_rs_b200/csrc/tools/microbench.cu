// microbench.cu — FP32 pipe micro-benchmarks for the SDF pair loop (sm_100a).
//
// Measures warp-instructions per clock per SM sub-partition for the instruction forms the pair
// loop uses, so the kernel's FP32 ceiling is known from measurement rather than from the
// datasheet: FFMA with uniform/constant operands vs three distinct vector registers, packed FFMA2
// with 1/2/3 register-pair operands, and the pair-loop body itself (scalar and packed) fed from
// registers.  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o microbench microbench.cu
#include <cstdio>
#include <cuda_runtime.h>
#include <vector>

#define ITERS 2048

template <int MODE> __global__ void __launch_bounds__(256) k(float *out, const float *in, float a, float b)
{
	const int tid = blockIdx.x * blockDim.x + threadIdx.x;
	float x[8], y[8], z[8];
#pragma unroll
	for (int i = 0; i < 8; ++i) {
		x[i] = in[(tid + i) & 1023];
		y[i] = in[(tid + 8 + i) & 1023];
		z[i] = in[(tid + 16 + i) & 1023];
	}
	for (int it = 0; it < ITERS; ++it) {
#pragma unroll
		for (int u = 0; u < 4; ++u) {
#pragma unroll
			for (int i = 0; i < 8; ++i) {
				if (MODE == 0) // FFMA reg, uniform, uniform
					x[i] = fmaf(x[i], a, b);
				else if (MODE == 1) // FFMA with three distinct vector registers
					x[i] = fmaf(x[i], y[i], z[i]);
				else if (MODE == 2) // FFMA accumulate form: d = y*z + d
					x[i] = fmaf(y[i], z[i], x[i]);
				else if (MODE == 3) // FFMA two regs + uniform
					x[i] = fmaf(x[i], y[i], b);
				else if (MODE == 4) // FFMA, the two other operands shared by all 8 chains (operand reuse cache)
					x[i] = fmaf(x[i], y[0], z[0]);
				else if (MODE == 5) // FFMA x = x*y + x (2 distinct registers)
					x[i] = fmaf(x[i], y[i], x[i]);
				else if (MODE == 6) // FFMA reg, reg, immediate
					x[i] = fmaf(x[i], y[i], 0.25f);
				else if (MODE == 7) // FFMA reg, immediate, reg
					x[i] = fmaf(x[i], 0.999f, z[i]);
				else if (MODE == 8) // FMUL reg, reg
					x[i] = x[i] * y[i];
				else if (MODE == 9) // FADD reg, reg
					x[i] = x[i] + y[i];
				else if (MODE == 10) // FMNMX reg, reg (ALU pipe)
					x[i] = fminf(x[i], y[i]);
				else if (MODE == 11) // FFMA.SAT three registers
					x[i] = __saturatef(fmaf(x[i], y[i], z[i]));
				else if (MODE == 12) // FFMA one shared operand: x = x*y0 + z_i
					x[i] = fmaf(x[i], y[0], z[i]);
			}
		}
	}
	float s = 0;
#pragma unroll
	for (int i = 0; i < 8; ++i)
		s += x[i] + y[i] + z[i];
	out[tid] = s;
}

template <int MODE> __global__ void __launch_bounds__(256) k2(float *out, const float *in, float a, float b)
{
	const int tid = blockIdx.x * blockDim.x + threadIdx.x;
	float2 x[8], y[8], z[8];
#pragma unroll
	for (int i = 0; i < 8; ++i) {
		x[i] = make_float2(in[(tid + i) & 1023], in[(tid + i + 3) & 1023]);
		y[i] = make_float2(in[(tid + 8 + i) & 1023], in[(tid + i + 5) & 1023]);
		z[i] = make_float2(in[(tid + 16 + i) & 1023], in[(tid + i + 7) & 1023]);
	}
	const float2 aa = make_float2(a, a), bb = make_float2(b, b);
	for (int it = 0; it < ITERS; ++it) {
#pragma unroll
		for (int u = 0; u < 4; ++u) {
#pragma unroll
			for (int i = 0; i < 8; ++i) {
				if (MODE == 0) // FFMA2 pair, uniform scalar, uniform scalar
					x[i] = __ffma2_rn(x[i], aa, bb);
				else if (MODE == 1) // FFMA2 three distinct pairs
					x[i] = __ffma2_rn(x[i], y[i], z[i]);
				else if (MODE == 2) // FFMA2 pair, register scalar (broadcast), pair
					x[i] = __ffma2_rn(x[i], make_float2(y[i].x, y[i].x), z[i]);
				else if (MODE == 3) // FFMA2 square-accumulate: d = y*y + d
					x[i] = __ffma2_rn(y[i], y[i], x[i]);
				else if (MODE == 4) // FMUL2 pair*pair
					x[i] = __fmul2_rn(x[i], y[i]);
				else if (MODE == 5) // FADD2 pair + register scalar
					x[i] = __fadd2_rn(x[i], make_float2(y[i].x, y[i].x));
			}
		}
	}
	float s = 0;
#pragma unroll
	for (int i = 0; i < 8; ++i)
		s += x[i].x + x[i].y + y[i].x + z[i].y;
	out[tid] = s;
}

// The pair-loop body, 16 pixels (4x4) against one segment per iteration, segment values in registers
// that change every iteration (so nothing is hoisted).  PACK as in sdf_kernel.cuh.
template <int PACK> __global__ void __launch_bounds__(128) body(float *out, const float *in, int iters)
{
	const int tid = blockIdx.x * blockDim.x + threadIdx.x;
	float mn[4][4];
#pragma unroll
	for (int r = 0; r < 4; ++r)
#pragma unroll
		for (int j = 0; j < 4; ++j)
			mn[r][j] = 1e30f;
	const float px0 = in[tid & 1023], py0 = in[(tid + 1) & 1023];
	float nvx = in[(tid + 2) & 1023], nvy = in[(tid + 3) & 1023], ndx = in[(tid + 4) & 1023], ndy = in[(tid + 5) & 1023];
	float dxn = in[(tid + 6) & 1023], dyn = in[(tid + 7) & 1023];
	const float step = in[(tid + 8) & 1023] * 1e-3f;
	for (int it = 0; it < iters; ++it) {
		nvx += step, nvy -= step, ndx += step, ndy -= step; // 4 extra FADD per segment (a real kernel has 2 LDS instead)
		if (PACK == 10) {
			// scalar, "operation-major": instructions that share two operands are emitted back to back
			// (t and qy row by row, qx column by column) so only d2 reads two fresh registers
			float pax[4], pay[4], t[4][4], m[4][4];
#pragma unroll
			for (int j = 0; j < 4; ++j)
				pax[j] = (px0 + (float)j) + nvx;
#pragma unroll
			for (int r = 0; r < 4; ++r) {
				pay[r] = (py0 + (float)r) + nvy;
				const float cr = pay[r] * dyn;
#pragma unroll
				for (int j = 0; j < 4; ++j)
					t[r][j] = __saturatef(fmaf(pax[j], dxn, cr));
#pragma unroll
				for (int j = 0; j < 4; ++j) {
					const float qy = fmaf(t[r][j], ndy, pay[r]);
					m[r][j] = qy * qy;
				}
			}
#pragma unroll
			for (int j = 0; j < 4; ++j)
#pragma unroll
				for (int r = 0; r < 4; ++r) {
					const float qx = fmaf(t[r][j], ndx, pax[j]);
					mn[r][j] = fminf(mn[r][j], fmaf(qx, qx, m[r][j]));
				}
		} else if (PACK == 0) {
			float pax[4];
#pragma unroll
			for (int j = 0; j < 4; ++j)
				pax[j] = (px0 + (float)j) + nvx;
#pragma unroll
			for (int r = 0; r < 4; ++r) {
				const float pay = (py0 + (float)r) + nvy;
				const float cr = pay * dyn;
#pragma unroll
				for (int j = 0; j < 4; ++j) {
					const float t = __saturatef(fmaf(pax[j], dxn, cr));
					const float qx = fmaf(t, ndx, pax[j]);
					const float qy = fmaf(t, ndy, pay);
					mn[r][j] = fminf(mn[r][j], fmaf(qx, qx, qy * qy));
				}
			}
		} else {
			float2 pax[2];
#pragma unroll
			for (int j = 0; j < 2; ++j)
				pax[j] = __fadd2_rn(make_float2(px0 + (float)(2 * j), px0 + (float)(2 * j + 1)), make_float2(nvx, nvx));
#pragma unroll
			for (int r = 0; r < 4; ++r) {
				const float pay1 = (py0 + (float)r) + nvy;
				const float2 pay = make_float2(pay1, pay1);
				const float cr = pay1 * dyn;
#pragma unroll
				for (int j = 0; j < 2; ++j) {
					float2 t;
					t.x = __saturatef(fmaf(pax[j].x, dxn, cr));
					t.y = __saturatef(fmaf(pax[j].y, dxn, cr));
					const float2 qx = __ffma2_rn(t, make_float2(ndx, ndx), pax[j]);
					const float2 qy = __ffma2_rn(t, make_float2(ndy, ndy), pay);
					float2 d2;
					if (PACK == 2)
						d2 = __ffma2_rn(qx, qx, __fmul2_rn(qy, qy));
					else {
						d2.x = fmaf(qx.x, qx.x, qy.x * qy.x);
						d2.y = fmaf(qx.y, qx.y, qy.y * qy.y);
					}
					mn[r][2 * j] = fminf(mn[r][2 * j], d2.x);
					mn[r][2 * j + 1] = fminf(mn[r][2 * j + 1], d2.y);
				}
			}
		}
	}
	float s = 0;
#pragma unroll
	for (int r = 0; r < 4; ++r)
#pragma unroll
		for (int j = 0; j < 4; ++j)
			s += mn[r][j];
	out[tid] = s;
}

static float *d_out, *d_in;
static int sms;
static double clock_ghz;

template <typename F> double time_ms(F launch)
{
	cudaEvent_t a, b;
	cudaEventCreate(&a);
	cudaEventCreate(&b);
	double best = 1e30;
	for (int r = 0; r < 5; ++r) {
		cudaEventRecord(a);
		launch();
		cudaEventRecord(b);
		cudaEventSynchronize(b);
		float ms;
		cudaEventElapsedTime(&ms, a, b);
		if (r > 0 && ms < best)
			best = ms;
	}
	return best;
}

int main()
{
	cudaDeviceProp p;
	cudaGetDeviceProperties(&p, 0);
	sms = p.multiProcessorCount;
	int khz = 0;
	cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
	clock_ghz = khz * 1e-6;
	printf("%s  SMs %d  max clock %.3f GHz (rates below assume it; real clock may be lower)\n", p.name, sms, clock_ghz);
	const int blocks = sms * 8;
	cudaMalloc(&d_out, (size_t)blocks * 256 * 4);
	std::vector<float> h(1024);
	for (int i = 0; i < 1024; ++i)
		h[i] = 0.5f + 0.001f * (float)(i % 97);
	cudaMalloc(&d_in, 4096);
	cudaMemcpy(d_in, h.data(), 4096, cudaMemcpyHostToDevice);

	auto report = [&](const char *name, double ms, double warp_instr_per_thread_iter, double flop_per_instr) {
		const double threads = (double)blocks * 256;
		const double instr = threads / 32 * warp_instr_per_thread_iter * ITERS;
		const double cyc = ms * 1e-3 * clock_ghz * 1e9;
		printf("%-44s %8.3f ms  %6.3f warp-instr/clk/SMSP  %7.2f TFLOP/s\n", name, ms, instr / cyc / (sms * 4),
		       instr * 32 * flop_per_instr / (ms * 1e-3) / 1e12);
	};
#define RUN1(M, NAME) report(NAME, time_ms([&] { k<M><<<blocks, 256>>>(d_out, d_in, 0.999f, 0.001f); }), 32, 2)
#define RUN2(M, NAME, FL) report(NAME, time_ms([&] { k2<M><<<blocks, 256>>>(d_out, d_in, 0.999f, 0.001f); }), 32, FL)
	RUN1(0, "FFMA  reg, uniform, uniform");
	RUN1(1, "FFMA  reg, reg, reg (3 distinct)");
	RUN1(2, "FFMA  d = y*z + d");
	RUN1(3, "FFMA  reg, reg, uniform");
	RUN1(4, "FFMA  x*y0+z0 (two operands shared by chains)");
	RUN1(12, "FFMA  x*y0+z_i (one operand shared)");
	RUN1(5, "FFMA  x*y+x (2 distinct regs)");
	RUN1(6, "FFMA  reg, reg, imm");
	RUN1(7, "FFMA  reg, imm, reg");
	RUN1(8, "FMUL  reg, reg");
	RUN1(9, "FADD  reg, reg");
	RUN1(10, "FMNMX reg, reg (ALU pipe)");
	RUN1(11, "FFMA.SAT reg, reg, reg");
	RUN2(0, "FFMA2 pair, uniform, uniform", 4);
	RUN2(1, "FFMA2 pair, pair, pair", 4);
	RUN2(2, "FFMA2 pair, scalar reg, pair", 4);
	RUN2(3, "FFMA2 d = y*y + d", 4);
	RUN2(4, "FMUL2 pair, pair", 2);
	RUN2(5, "FADD2 pair, scalar reg", 2);

	const int it = 4096;
	const int bblocks = sms * 16;
	auto body_report = [&](const char *name, double ms) {
		const double pairs = (double)bblocks * 128 * 16 * it;
		const double cyc = ms * 1e-3 * clock_ghz * 1e9;
		printf("%-44s %8.3f ms  %6.2f clk per warp-pair/SMSP  %7.2f TFLOP/s (11 flop/pair)  %.3e pairs/s\n", name, ms,
		       cyc * sms * 4 / (pairs / 32), pairs * 11 / (ms * 1e-3) / 1e12, pairs / (ms * 1e-3));
	};
	body_report("pair-loop body, scalar (PACK 0)", time_ms([&] { body<0><<<bblocks, 128>>>(d_out, d_in, it); }));
	body_report("pair-loop body, scalar operation-major", time_ms([&] { body<10><<<bblocks, 128>>>(d_out, d_in, it); }));
	body_report("pair-loop body, FFMA2 projection (PACK 1)", time_ms([&] { body<1><<<bblocks, 128>>>(d_out, d_in, it); }));
	body_report("pair-loop body, FFMA2 throughout (PACK 2)", time_ms([&] { body<2><<<bblocks, 128>>>(d_out, d_in, it); }));
	return 0;
}
