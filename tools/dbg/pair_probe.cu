// Debug probe (run on a GPU box): the packed-pair gradient columns (rnea_grad_col2_rt) against the scalar columns (rnea_grad_col_rt) ON THE DEVICE, mismatches
// per column and output joint.  This is the reproducer of the ptxas 12.9 defect described at rbd.cuh: mul2 (mul.rn.f32x2 + add.rn.f32x2 fused into FFMA2);
// -DGATO_F2_EMULATE=<bits> evaluates fma2 / add2 / mul2 lane by lane to bisect.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -fmad=false --expt-relaxed-constexpr -o tools/dbg/pair_probe tools/dbg/pair_probe.cu
#include <cstdio>
#include <vector>
#include <random>
#include "../../gato_b200/csrc/items.cuh"
using namespace gato;
template<class P, int W>
__global__ void k_probe(const float* xux_all, float* out_s, float* out_p, int n)
{
        constexpr int NQ = P::NQ;
        const int     i = blockIdx.x * blockDim.x + threadIdx.x;
        if (i >= n) return;
        float xux[5 * NQ], fext[6] = {0, 0, 0, 0, 0, 0};
        for (int e = 0; e < 5 * NQ; e++) xux[e] = xux_all[i * 5 * NQ + e];
        typename Rbd<P>::DynState st;
        Rbd<P>::dyn_prologue(xux, xux + NQ, xux + 2 * NQ, fext, st);
        for (int k = 0; k < NQ; k++) {
                float dc[NQ];
                Rbd<P>::template rnea_grad_col_rt<W>(k, st.X, xux + NQ, st.v, st.a, st.f, st.Iv, st.FxvI, dc);
                for (int j = 0; j < NQ; j++) out_s[(i * NQ + k) * NQ + j] = dc[j];
        }
        for (int k0 = 0; k0 < NQ; k0 += 2) {
                f2 dc[NQ];
                Rbd<P>::template rnea_grad_col2_rt<W>(k0, st.X, xux + NQ, st.v, st.a, st.f, st.Iv, st.FxvI, dc);
                for (int j = 0; j < NQ; j++) {
                        out_p[(i * NQ + k0) * NQ + j] = dc[j].x;
                        if (k0 + 1 < NQ) out_p[(i * NQ + k0 + 1) * NQ + j] = dc[j].y;
                }
        }
}
template<class P, int W>
void run(const char* name)
{
        constexpr int NQ = P::NQ;
        const int     n = 64;
        std::mt19937                          g(3);
        std::uniform_real_distribution<float> d(-1.5f, 1.5f);
        std::vector<float>                    h(n * 5 * NQ);
        for (auto& x : h) x = d(g);
        float *dx, *ds, *dp;
        cudaMalloc(&dx, h.size() * 4), cudaMalloc(&ds, n * NQ * NQ * 4), cudaMalloc(&dp, n * NQ * NQ * 4);
        cudaMemcpy(dx, h.data(), h.size() * 4, cudaMemcpyHostToDevice);
        k_probe<P, W><<<2, 32>>>(dx, ds, dp, n);
        std::vector<float> s(n * NQ * NQ), p(n * NQ * NQ);
        cudaMemcpy(s.data(), ds, s.size() * 4, cudaMemcpyDeviceToHost), cudaMemcpy(p.data(), dp, p.size() * 4, cudaMemcpyDeviceToHost);
        printf("%s W=%d err=%s\n", name, W, cudaGetErrorString(cudaGetLastError()));
        int cnt[NQ][NQ] = {};
        for (int i = 0; i < n; i++)
                for (int k = 0; k < NQ; k++)
                        for (int j = 0; j < NQ; j++)
                                if (s[(i * NQ + k) * NQ + j] != p[(i * NQ + k) * NQ + j]) cnt[k][j]++;
        for (int k = 0; k < NQ; k++) {
                printf("  column %d mismatches per output joint:", k);
                for (int j = 0; j < NQ; j++) printf(" %d", cnt[k][j]);
                printf("\n");
        }
}
int main()
{
        run<Iiwa14, 0>("iiwa14");
        run<Iiwa14, 1>("iiwa14");
        run<Indy7, 0>("indy7");
        return 0;
}
