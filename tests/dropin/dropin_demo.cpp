// dropin_demo.cpp -- a host program written against the REFERENCE's headers
// (guided_filter.h / guided_filter_d.h), exercising the same call sequences as its demos:
//   cudaGuidedFilterDemo  (GuidedFilter/main.cpp:132-150): cudaMallocPitch, init(w,h,gch,sch), run
//   cudaSmallGuidedDemo   (GuidedFilter/main.cpp:221-267): hGuidedFilter with A/B scratch
// It is compiled against include/ and linked with libgf_b200.so to show the library is a
// drop-in at the source level.  Usage: dropin_demo <in.bin> <out.bin>
//   in.bin : int32 w, h, r, sch; float eps; float guide[h*w]; float src[h*w*sch]
//   out.bin: float q_class[h*w*sch] (GuidedFilter::run); float q_fused[h*w], A[h*w], B[h*w]
//            (hGuidedFilter on channel 0 of src); float q_chain[h*w*sch] (the eleven launcher calls of
//            GuidedFilter::run, guided_filter.cpp:28-66, made by hand: hBoxFilter, hMultiply, hCalcA, hCalcB,
//            hLinearTransform on cudaMallocPitch planes with the reference's int4 whcs descriptors)
#include <cstdio>
#include <memory>
#include <vector>

#include "guided_filter.h"
#include "guided_filter_d.h"

int main(int argc, char** argv)
{
    if (argc < 3) return 2;
    if (!initDevice(0)) return 3;
    FILE* f = std::fopen(argv[1], "rb");
    if (!f) return 4;
    int w = 0, h = 0, r = 0, sch = 0;
    float eps = 0;
    if (std::fread(&w, 4, 1, f) != 1 || std::fread(&h, 4, 1, f) != 1 || std::fread(&r, 4, 1, f) != 1 ||
        std::fread(&sch, 4, 1, f) != 1 || std::fread(&eps, 4, 1, f) != 1) return 5;
    std::vector<float> guide((size_t)w * h), src((size_t)w * h * sch);
    if (std::fread(guide.data(), 4, guide.size(), f) != guide.size() || std::fread(src.data(), 4, src.size(), f) != src.size()) return 5;
    std::fclose(f);

    // --- path A call sequence (main.cpp:126-150)
    float *d_src = nullptr, *d_guidiance = nullptr, *d_dst = nullptr;
    const size_t spitch1 = w * sch * sizeof(float), gpitch1 = w * sizeof(float);
    size_t spitch2 = 0, gpitch2 = 0;
    CHECK(cudaMallocPitch(reinterpret_cast<void**>(&d_src), &spitch2, spitch1, h));
    CHECK(cudaMallocPitch(reinterpret_cast<void**>(&d_guidiance), &gpitch2, gpitch1, h));
    CHECK(cudaMallocPitch(reinterpret_cast<void**>(&d_dst), &spitch2, spitch1, h));
    CHECK(cudaMemcpy2D(d_src, spitch2, src.data(), spitch1, spitch1, h, cudaMemcpyHostToDevice));
    CHECK(cudaMemcpy2D(d_guidiance, gpitch2, guide.data(), gpitch1, gpitch1, h, cudaMemcpyHostToDevice));
    std::shared_ptr<GuidedFilter> gfilter = std::make_shared<GuidedFilter>();
    gfilter->init(w, h, 1, sch);
    GpuTimer timer(0);
    gfilter->run(d_guidiance, d_src, d_dst, r, eps);
    float t_elapsed = timer.read();
    std::printf("Time of cuda guided filter: %fms\n", t_elapsed);
    std::vector<float> q_class((size_t)w * h * sch);
    CHECK(cudaMemcpy2D(q_class.data(), spitch1, d_dst, spitch2, spitch1, h, cudaMemcpyDeviceToHost));

    // --- the body of GuidedFilter::run (guided_filter.cpp:57-65) launcher by launcher, scratch planes as in
    //     allocMemory (guided_filter.cpp:70-93): int4 = {.x height, .y channels, .z stride, .w width}
    std::vector<float> q_chain((size_t)w * h * sch);
    {
        int4 swhcs, gwhcs, iwhcs;
        swhcs.w = w; swhcs.x = h; swhcs.y = sch; swhcs.z = static_cast<int>(spitch2 / sizeof(float));
        gwhcs.w = w; gwhcs.x = h; gwhcs.y = 1; gwhcs.z = static_cast<int>(gpitch2 / sizeof(float));
        iwhcs.w = w + 1; iwhcs.x = h + 1; iwhcs.y = sch; iwhcs.z = 0;
        float *pm, *im, *ipm, *iim, *a, *b, *am, *bm, *d_chain;
        size_t ps = 0, pg = 0;
        float** splanes[] = {&pm, &ipm, &a, &b, &am, &bm, &d_chain};
        for (float** pp : splanes) CHECK(cudaMallocPitch(reinterpret_cast<void**>(pp), &ps, spitch1, h));
        CHECK(cudaMallocPitch(reinterpret_cast<void**>(&im), &pg, gpitch1, h));
        CHECK(cudaMallocPitch(reinterpret_cast<void**>(&iim), &pg, gpitch1, h));
        if (ps != spitch2 || pg != gpitch2) return 7;
        hBoxFilter(d_src, pm, nullptr, swhcs, iwhcs, r);
        hBoxFilter(d_guidiance, im, nullptr, gwhcs, iwhcs, r);
        hMultiply(d_src, d_guidiance, ipm, swhcs, gwhcs);
        hMultiply(d_guidiance, d_guidiance, iim, gwhcs, gwhcs);
        hBoxFilter(ipm, ipm, nullptr, swhcs, iwhcs, r);
        hBoxFilter(iim, iim, nullptr, gwhcs, iwhcs, r);
        hCalcA(a, pm, im, ipm, iim, swhcs, gwhcs, eps);
        hCalcB(b, a, pm, im, swhcs, gwhcs);
        hBoxFilter(a, am, nullptr, swhcs, iwhcs, r);
        hBoxFilter(b, bm, nullptr, swhcs, iwhcs, r);
        hLinearTransform(d_guidiance, d_chain, am, bm, gwhcs, swhcs);
        CHECK(cudaDeviceSynchronize());
        CHECK(cudaMemcpy2D(q_chain.data(), spitch1, d_chain, spitch2, spitch1, h, cudaMemcpyDeviceToHost));
        for (float** pp : splanes) CUDA_SAFE_FREE(*pp);
        CUDA_SAFE_FREE(im); CUDA_SAFE_FREE(iim);
    }

    // --- path B call sequence (main.cpp:215-279) on channel 0 of src
    std::vector<float> src0((size_t)w * h);
    for (size_t i = 0; i < src0.size(); ++i) src0[i] = src[i * sch];
    float *d_s = nullptr, *d_g = nullptr, *d_q = nullptr, *d_A = nullptr, *d_B = nullptr;
    size_t spitch = w * sizeof(float), dpitch = 0;
    CHECK(cudaMallocPitch(reinterpret_cast<void**>(&d_s), &dpitch, spitch, h));
    CHECK(cudaMallocPitch(reinterpret_cast<void**>(&d_g), &dpitch, spitch, h));
    CHECK(cudaMallocPitch(reinterpret_cast<void**>(&d_q), &dpitch, spitch, h));
    CHECK(cudaMallocPitch(reinterpret_cast<void**>(&d_A), &dpitch, spitch, h));
    CHECK(cudaMallocPitch(reinterpret_cast<void**>(&d_B), &dpitch, spitch, h));
    const int stride = static_cast<int>(dpitch / sizeof(float));
    CHECK(cudaMemcpy2D(d_s, dpitch, src0.data(), spitch, spitch, h, cudaMemcpyHostToDevice));
    CHECK(cudaMemcpy2D(d_g, dpitch, guide.data(), spitch, spitch, h, cudaMemcpyHostToDevice));
    hGuidedFilter(d_g, d_s, d_q, d_A, d_B, eps, r, w, h, stride);
    CHECK(cudaDeviceSynchronize());
    std::vector<float> q((size_t)w * h), A((size_t)w * h), B((size_t)w * h);
    CHECK(cudaMemcpy2D(q.data(), spitch, d_q, dpitch, spitch, h, cudaMemcpyDeviceToHost));
    CHECK(cudaMemcpy2D(A.data(), spitch, d_A, dpitch, spitch, h, cudaMemcpyDeviceToHost));
    CHECK(cudaMemcpy2D(B.data(), spitch, d_B, dpitch, spitch, h, cudaMemcpyDeviceToHost));

    f = std::fopen(argv[2], "wb");
    if (!f) return 6;
    std::fwrite(q_class.data(), 4, q_class.size(), f);
    std::fwrite(q.data(), 4, q.size(), f);
    std::fwrite(A.data(), 4, A.size(), f);
    std::fwrite(B.data(), 4, B.size(), f);
    std::fwrite(q_chain.data(), 4, q_chain.size(), f);
    std::fclose(f);
    CUDA_SAFE_FREE(d_src); CUDA_SAFE_FREE(d_guidiance); CUDA_SAFE_FREE(d_dst);
    CUDA_SAFE_FREE(d_s); CUDA_SAFE_FREE(d_g); CUDA_SAFE_FREE(d_q); CUDA_SAFE_FREE(d_A); CUDA_SAFE_FREE(d_B);
    return 0;
}
