// gko::acc::row_major<ValueType, Dimensionality> -- the plain accessor: no
// precision change between storage and arithmetic, element access returns a
// reference.  Clean-room (see accessor_config.hpp); the reference itself only
// instantiates reduced_row_major, its README (README.md:19) says "all other
// accessors work accordingly" -- which holds here because the kernel-level
// templates (accblas/{gemv,dot,trsv}_kernels.cuh) only use `range(i, j)`,
// `range.length(d)` and `accessor::arithmetic_type`.  Template parameter
// order and names follow upstream Ginkgo from memory; not verifiable offline.
#pragma once

#include <array>
#include <cstddef>
#include <type_traits>

#include "accessor_config.hpp"

namespace gko {
namespace acc {

template <typename ValueType, std::size_t Dimensionality>
class row_major {
public:
    using value_type = ValueType;
    using arithmetic_type = typename std::remove_cv<ValueType>::type;
    using storage_type = ValueType;
    static constexpr size_type dimensionality =
        static_cast<size_type>(Dimensionality);
    using const_accessor = row_major<const ValueType, Dimensionality>;
    using length_type = std::array<size_type, Dimensionality>;
    using stride_type =
        std::array<size_type, (Dimensionality > 0 ? Dimensionality - 1 : 0)>;

    static_assert(Dimensionality >= 1, "at least one dimension is required");

    template <typename SizeArray, typename StrideArray>
    constexpr GKO_ACC_ATTRIBUTES row_major(const SizeArray& size,
                                           value_type* data,
                                           const StrideArray& stride)
        : size_{}, stride_{}, data_(data)
    {
        for (std::size_t d = 0; d < Dimensionality; ++d) {
            size_[d] = static_cast<size_type>(size[d]);
        }
        for (std::size_t d = 0; d + 1 < Dimensionality; ++d) {
            stride_[d] = static_cast<size_type>(stride[d]);
        }
    }

    template <typename SizeArray>
    constexpr GKO_ACC_ATTRIBUTES row_major(const SizeArray& size,
                                           value_type* data)
        : size_{}, stride_{}, data_(data)
    {
        for (std::size_t d = 0; d < Dimensionality; ++d) {
            size_[d] = static_cast<size_type>(size[d]);
        }
        size_type run = 1;
        for (std::size_t d = Dimensionality - 1; d > 0; --d) {
            run *= size_[d];
            stride_[d - 1] = run;
        }
    }

    // the last index is contiguous, index d < D-1 is scaled by stride[d]
    template <typename... Indices>
    constexpr GKO_ACC_ATTRIBUTES GKO_ACC_INLINE value_type& operator()(
        Indices... indices) const
    {
        static_assert(sizeof...(Indices) == Dimensionality,
                      "one index per dimension");
        const size_type idx[Dimensionality] = {
            static_cast<size_type>(indices)...};
        size_type offset = idx[Dimensionality - 1];
        for (std::size_t d = 0; d + 1 < Dimensionality; ++d) {
            offset += idx[d] * stride_[d];
        }
        return data_[offset];
    }

    constexpr GKO_ACC_ATTRIBUTES GKO_ACC_INLINE size_type
    length(size_type dim) const
    {
        return dim < dimensionality ? size_[dim] : size_type{1};
    }

    constexpr GKO_ACC_ATTRIBUTES const length_type& get_size() const
    {
        return size_;
    }
    constexpr GKO_ACC_ATTRIBUTES const stride_type& get_stride() const
    {
        return stride_;
    }
    constexpr GKO_ACC_ATTRIBUTES value_type* get_stored_data() const
    {
        return data_;
    }
    constexpr GKO_ACC_ATTRIBUTES const arithmetic_type* get_const_storage()
        const
    {
        return data_;
    }
    constexpr GKO_ACC_ATTRIBUTES const_accessor to_const() const
    {
        return const_accessor{size_, data_, stride_};
    }

private:
    length_type size_;
    stride_type stride_;
    value_type* data_;
};

}  // namespace acc
}  // namespace gko
