// Shared pieces of the three benchmark drivers.  The drivers print the same
// CSV the reference's drivers print (/root/reference/cuda/*_benchmark.cu:
// ';' delimiter, precision(16), scientific, showpos for GEMV/TRSV) so existing
// plotting scripts keep working, but the fixtures are generated and converted
// ON THE DEVICE (accblas_fill_uniform / accblas_convert) from the same
// std::default_random_engine(42) + uniform(-1,1) stream, in the same draw
// order as the reference's *Memory classes.
#pragma once

#include <cstdint>
#include <functional>
#include <iostream>
#include <string>
#include <tuple>
#include <vector>

#include <accblas/utils.cuh>

namespace driver {

constexpr char DELIM{';'};
constexpr std::uint32_t seed{42};

template <typename T>
class DeviceBuffer {
public:
    explicit DeviceBuffer(std::int64_t count) : count_(count)
    {
        CUDA_CALL(cudaMalloc(&ptr_, sizeof(T) * static_cast<std::size_t>(
                                                    count > 0 ? count : 1)));
    }
    DeviceBuffer(const DeviceBuffer&) = delete;
    DeviceBuffer& operator=(const DeviceBuffer&) = delete;
    ~DeviceBuffer()
    {
        cudaDeviceSynchronize();
        cudaFree(ptr_);
    }
    T* data() { return ptr_; }
    const T* data() const { return ptr_; }
    std::int64_t size() const { return count_; }

    void copy_from(const DeviceBuffer& other)
    {
        CUDA_CALL(cudaMemcpy(ptr_, other.ptr_, sizeof(T) * count_,
                             cudaMemcpyDeviceToDevice));
    }
    std::vector<T> to_host(std::int64_t count) const
    {
        std::vector<T> out(static_cast<std::size_t>(count));
        CUDA_CALL(cudaMemcpy(out.data(), ptr_, sizeof(T) * count,
                             cudaMemcpyDeviceToHost));
        return out;
    }

private:
    T* ptr_{nullptr};
    std::int64_t count_;
};

// out = static_cast<Dst>(in), element-wise on the device
template <typename Dst, typename Src>
void convert(accblas_handle_t h, std::int64_t rows, std::int64_t cols,
             std::int64_t ld, const DeviceBuffer<Src>& in, DeviceBuffer<Dst>& out)
{
    ACCBLAS_CALL(accblas_convert(h, accblas_detail::dtype_of<Dst>::value,
                                 accblas_detail::dtype_of<Src>::value, rows,
                                 cols, in.data(), ld, out.data(), ld, nullptr));
}

// rows x cols block of the global stream starting at draw `first`
template <typename T>
void fill(accblas_handle_t h, std::int64_t rows, std::int64_t cols,
          std::int64_t ld, std::uint64_t first, DeviceBuffer<T>& out)
{
    ACCBLAS_CALL(accblas_fill_uniform(h, accblas_detail::dtype_of<T>::value,
                                      rows, cols, out.data(), ld, seed, first,
                                      nullptr));
}

// storage value -> double on the host (exact)
inline double widen(double v) { return v; }
inline double widen(float v) { return static_cast<double>(v); }
inline double widen(__half v) { return static_cast<double>(__half2float(v)); }

struct Options {
    bool measure_error{false};
    bool fp16{false};       // extension: append the fp16-storage variants
    bool only_max{false};   // extension: run the maximum size only
    std::int64_t max_size{0};
};

inline bool parse(int argc, char** argv, std::int64_t default_max,
                  std::int64_t min_size, const char* what, Options& opt)
{
    opt.max_size = default_max;
    const std::string err_flag("--error"), size_flag("--size"),
        fp16_flag("--fp16"), exact_flag("--exact");
    auto usage = [&]() {
        std::cerr << "Usage: " << argv[0] << " [" << err_flag << "] ["
                  << size_flag << "=SIZE] [" << fp16_flag << "] ["
                  << exact_flag << "]\n"
                  << "With:\n"
                  << err_flag << ":    compute errors of the " << what << "\n"
                  << size_flag << ":     set the maximum size. Default value: "
                  << default_max << "; Min value: " << min_size << '\n'
                  << fp16_flag
                  << ":     also run the fp16-storage accessor variants\n"
                  << exact_flag << ":    run SIZE only instead of the sweep\n"
                  << "Without parameters: benchmark different " << what << '\n';
    };
    for (int i = 1; i < argc; ++i) {
        const std::string cur(argv[i]);
        if (cur == err_flag) {
            opt.measure_error = true;
        } else if (cur == fp16_flag) {
            opt.fp16 = true;
        } else if (cur == exact_flag) {
            opt.only_max = true;
        } else if (cur.substr(0, size_flag.size()) == size_flag &&
                   cur.size() > size_flag.size() + 1) {
            opt.max_size = std::stoll(cur.substr(size_flag.size() + 1));
        } else {
            std::cerr << "Unsupported parameter: " << cur << '\n';
            usage();
            return false;
        }
    }
    if (opt.max_size < min_size) {
        std::cerr << "The size needs to be at least " << min_size << '\n';
        return false;
    }
    return true;
}

}  // namespace driver
