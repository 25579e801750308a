// tools/pipe_bench.cu -- issue-rate microbenchmarks for instruction mixes on sm_100a (scratch tool, not product).
// Each mix is a string over {L: LOP3, I: IMAD, H: IMAD.HI, W: IMAD.WIDE, S: SHF, A: IADD3, P: POPC, D: DP4A, F: FFMA, M: FMNMX/VIMNMX}
// executed round-robin on 16 independent accumulators; prints lane-ops/s per letter class and total.
#include <cstdio>
#include <cstdint>
#include <cstring>
#include <cuda_runtime.h>

template <char C>
__device__ __forceinline__ void op(uint32_t &a, uint64_t &w, uint32_t b, uint32_t c) {
    if (C == 'L') asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(a) : "r"(b), "r"(c));
    else if (C == 'I') asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(a) : "r"(b), "r"(c));
    else if (C == 'H') asm volatile("mad.hi.u32 %0, %0, %1, %2;" : "+r"(a) : "r"(b), "r"(c));
    else if (C == 'W') { uint32_t lo = (uint32_t)w; asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(w) : "r"(lo), "r"(b)); }
    else if (C == 'S') asm volatile("shf.l.wrap.b32 %0, %0, %1, 1;" : "+r"(a) : "r"(b));
    else if (C == 'A') asm volatile("add.u32 %0, %0, %1;" : "+r"(a) : "r"(b));
    else if (C == 'P') asm volatile("popc.b32 %0, %0;" : "+r"(a));
    else if (C == 'D') asm volatile("dp4a.u32.u32 %0, %0, %1, %2;" : "+r"(a) : "r"(b), "r"(c));
    else if (C == 'F') asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+r"(a) : "r"(b), "r"(c));
    else if (C == 'M') asm volatile("min.u32 %0, %0, %1;" : "+r"(a) : "r"(b));
    else if (C == 'J') asm volatile("mad.lo.u32 %0, %1, 0xffffffff, %0;" : "+r"(a) : "r"(b));   // a - b as IMAD imm
    else if (C == 'K') asm volatile("mad.lo.u32 %0, %0, 3, %1;" : "+r"(a) : "r"(b));           // IMAD imm, 2 reg reads
    else if (C == 'l') asm volatile("lop3.b32 %0, %0, %1, 0x55555555, 0x96;" : "+r"(a) : "r"(b)); // LOP3 with 2 reg reads
    else if (C == 'X') asm volatile("shr.u32 %0, %0, 31;" : "+r"(a));
}

template <char... Cs> struct Mix {
    static constexpr int N = sizeof...(Cs);
    template <int K> static __device__ __forceinline__ void run(uint32_t (&a)[16], uint64_t (&w)[16], uint32_t b, uint32_t c) {
        int k = K;
        ((op<Cs>(a[k & 15], w[k & 15], b, c), ++k), ...);
    }
};

template <typename M>
__global__ void __launch_bounds__(256) bench(uint32_t *out, int iters, uint32_t b, uint32_t c) {
    constexpr int REPS = (M::N >= 8) ? 8 : 16;   // keep the loop body inside the instruction caches
    uint32_t a[16];
    uint64_t w[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) { a[i] = threadIdx.x + i * 7919u; w[i] = a[i]; }
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int rep = 0; rep < REPS; ++rep) M::template run<0>(a, w, b, c);
    }
    uint32_t x = 0;
#pragma unroll
    for (int i = 0; i < 16; ++i) x ^= a[i] ^ (uint32_t)w[i] ^ (uint32_t)(w[i] >> 32);
    out[blockIdx.x * blockDim.x + threadIdx.x] = x;
}

template <typename M>
void run(const char *name, uint32_t *d_out, int sms) {
    const int grid = sms * 8, iters = 4096;
    const int REPS = (M::N >= 8) ? 8 : 16;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    float best = 1e30f;
    for (int r = 0; r < 4; ++r) {
        cudaEventRecord(e0);
        bench<M><<<grid, 256>>>(d_out, iters, 0x9E3779B9u, 2u);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        if (r) best = ms < best ? ms : best;
    }
    const double ops = (double)grid * 256 * iters * (double)REPS * M::N;
    const double warp_instr_per_smsp = (double)grid * 8 / (sms * 4.0) * iters * (double)REPS * M::N;
    const double cyc = best * 1e-3 * 1.965e9;
    printf("{\"mix\": \"%s\", \"n\": %d, \"Tops\": %.3f, \"cycles_per_mix_per_smsp_warp\": %.3f}\n", name, M::N, ops / (best * 1e-3) / 1e12,
           cyc / (warp_instr_per_smsp / M::N));
    fflush(stdout);
}

int main() {
    cudaDeviceProp prop; cudaGetDeviceProperties(&prop, 0);
    uint32_t *d; cudaMalloc(&d, (size_t)prop.multiProcessorCount * 8 * 256 * 4);
    const int s = prop.multiProcessorCount;
#define R(name, ...) run<Mix<__VA_ARGS__>>(name, d, s)
    R("L", 'L'); R("I", 'I'); R("H", 'H'); R("W", 'W'); R("S", 'S'); R("A", 'A'); R("P", 'P'); R("D", 'D'); R("F", 'F'); R("M", 'M'); R("X", 'X');
    R("LI", 'L', 'I'); R("LLI", 'L', 'L', 'I'); R("LLLLI", 'L', 'L', 'L', 'L', 'I');
    R("LH", 'L', 'H'); R("LLH", 'L', 'L', 'H'); R("LLLLH", 'L', 'L', 'L', 'L', 'H');
    R("LW", 'L', 'W'); R("LLW", 'L', 'L', 'W'); R("LLLLW", 'L', 'L', 'L', 'L', 'W');
    R("LLLLD", 'L', 'L', 'L', 'L', 'D'); R("LLD", 'L', 'L', 'D');
    R("LLLLP", 'L', 'L', 'L', 'L', 'P'); R("LLLLLLLLP", 'L', 'L', 'L', 'L', 'L', 'L', 'L', 'L', 'P');
    R("LLLLF", 'L', 'L', 'L', 'L', 'F'); R("LF", 'L', 'F');
    R("word_orig_8L2S", 'L', 'L', 'L', 'L', 'S', 'L', 'L', 'L', 'L', 'S');
    R("word_8L2I2H", 'L', 'L', 'I', 'L', 'H', 'L', 'L', 'I', 'L', 'L', 'H', 'L');
    R("word_8L2W2I", 'L', 'L', 'W', 'L', 'I', 'L', 'L', 'W', 'L', 'L', 'I', 'L');
    R("word_9L2W", 'L', 'L', 'W', 'L', 'L', 'L', 'L', 'W', 'L', 'L', 'L');
    R("word_8L4I", 'L', 'L', 'I', 'L', 'I', 'L', 'L', 'I', 'L', 'L', 'I', 'L');
    R("word_8L6I", 'L', 'I', 'L', 'I', 'L', 'I', 'L', 'I', 'L', 'I', 'L', 'I', 'L', 'L');
    R("J", 'J'); R("K", 'K'); R("l", 'l');
    R("LJ", 'L', 'J'); R("LK", 'L', 'K'); R("lK", 'l', 'K');
    R("word_8L6K", 'L', 'K', 'L', 'K', 'L', 'K', 'L', 'K', 'L', 'K', 'L', 'K', 'L', 'L');
    R("word_8l6K", 'l', 'K', 'l', 'K', 'l', 'K', 'l', 'K', 'l', 'K', 'l', 'K', 'l', 'l');
    R("word_8L6J", 'L', 'J', 'L', 'J', 'L', 'J', 'L', 'J', 'L', 'J', 'L', 'J', 'L', 'L');
    R("word_8L4K", 'L', 'L', 'K', 'L', 'K', 'L', 'L', 'K', 'L', 'L', 'K', 'L');
    R("word_8L8K", 'L', 'K', 'L', 'K', 'L', 'K', 'L', 'K', 'L', 'K', 'L', 'K', 'L', 'K', 'L', 'K');
    R("word_8L", 'L', 'L', 'L', 'L', 'L', 'L', 'L', 'L');
    R("word_8L4D", 'L', 'L', 'D', 'L', 'D', 'L', 'L', 'D', 'L', 'L', 'D', 'L');
    R("word_8L4F", 'L', 'L', 'F', 'L', 'F', 'L', 'L', 'F', 'L', 'L', 'F', 'L');
    return 0;
}
