// TEST INFRASTRUCTURE ONLY.  Host-only program (no GPU needed) that pins the
// software binary16 conversions of oracle_common.hpp against CUDA's own
// cuda_fp16.h host implementations -- the functions static_cast<__half>(double)
// / static_cast<__half>(float) / operator float() resolve to.
// Usage: fp16_pin <in.f64> <n> <out_from_f64.u16> <out_from_f32.u16> <out_back.f32>
#include <cuda_fp16.h>

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

int main(int argc, char** argv)
{
    if (argc != 6) {
        return 2;
    }
    const long n = atol(argv[2]);
    std::vector<double> in(n);
    FILE* f = fopen(argv[1], "rb");
    if (!f || fread(in.data(), sizeof(double), n, f) != (size_t)n) {
        return 3;
    }
    fclose(f);
    std::vector<unsigned short> from64(n), from32(n);
    std::vector<float> back(n);
    for (long i = 0; i < n; ++i) {
        const __half a = __double2half(in[i]);
        const __half b = __float2half_rn(static_cast<float>(in[i]));
        from64[i] = __half_as_ushort(a);
        from32[i] = __half_as_ushort(b);
        back[i] = __half2float(a);
    }
    f = fopen(argv[3], "wb");
    fwrite(from64.data(), 2, n, f);
    fclose(f);
    f = fopen(argv[4], "wb");
    fwrite(from32.data(), 2, n, f);
    fclose(f);
    f = fopen(argv[5], "wb");
    fwrite(back.data(), 4, n, f);
    fclose(f);
    return 0;
}
