// dropin_demo.cpp -- a host program written against the REFERENCE's headers
// (guided_filter.h / guided_filter_d.h), exercising the same call sequences as its demos:
//   cudaGuidedFilterDemo  (GuidedFilter/main.cpp:132-150): cudaMallocPitch, init(w,h,gch,sch), run
//   cudaSmallGuidedDemo   (GuidedFilter/main.cpp:221-267): hGuidedFilter with A/B scratch
// It is compiled against include/ and linked with libgf_b200.so to show the library is a
// drop-in at the source level.  Usage: dropin_demo <in.bin> <out.bin>
//   in.bin : int32 w, h, r, sch; float eps; float guide[h*w]; float src[h*w*sch]
//   out.bin: float q_class[h*w*sch] (GuidedFilter::run); float q_fused[h*w], A[h*w], B[h*w]
//            (hGuidedFilter on channel 0 of src)
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
    std::fclose(f);
    CUDA_SAFE_FREE(d_src); CUDA_SAFE_FREE(d_guidiance); CUDA_SAFE_FREE(d_dst);
    CUDA_SAFE_FREE(d_s); CUDA_SAFE_FREE(d_g); CUDA_SAFE_FREE(d_q); CUDA_SAFE_FREE(d_A); CUDA_SAFE_FREE(d_B);
    return 0;
}
