// gko::acc::block_col_major<ValueType, Dimensionality> -- clean-room (see
// accessor_config.hpp for provenance; interface from memory of upstream
// Ginkgo, NOT verifiable offline: the reference never instantiates it).
//
// A stack of column-major blocks: the LAST TWO indices address a block in
// column-major order (row index contiguous), the indices in front of them are
// row-major over blocks.  For Dimensionality == 2 this is a plain column-major
// matrix.  No precision change: element access returns a reference.
//     offset(i0, ..., r, c) = sum_d i_d * stride[d] + c * stride[D-2] + r
#pragma once

#include <array>
#include <cstddef>
#include <type_traits>

#include "accessor_config.hpp"

namespace gko {
namespace acc {

template <typename ValueType, std::size_t Dimensionality>
class block_col_major {
public:
    using value_type = ValueType;
    using arithmetic_type = typename std::remove_cv<ValueType>::type;
    using storage_type = ValueType;
    static constexpr size_type dimensionality =
        static_cast<size_type>(Dimensionality);
    using const_accessor = block_col_major<const ValueType, Dimensionality>;
    using length_type = std::array<size_type, Dimensionality>;
    using stride_type = std::array<size_type, Dimensionality - 1>;

    static_assert(Dimensionality >= 2,
                  "a block needs a row and a column index");

    // stride[D-2] = distance between columns of a block; stride[d < D-2] =
    // distance between consecutive values of index d
    template <typename SizeArray, typename StrideArray>
    constexpr GKO_ACC_ATTRIBUTES block_col_major(const SizeArray& size,
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

    // dense: columns of `rows` values, blocks of rows * cols values, ...
    template <typename SizeArray>
    constexpr GKO_ACC_ATTRIBUTES block_col_major(const SizeArray& size,
                                                 value_type* data)
        : size_{}, stride_{}, data_(data)
    {
        for (std::size_t d = 0; d < Dimensionality; ++d) {
            size_[d] = static_cast<size_type>(size[d]);
        }
        stride_[Dimensionality - 2] = size_[Dimensionality - 2];
        size_type run = size_[Dimensionality - 2] * size_[Dimensionality - 1];
        for (std::size_t d = Dimensionality - 2; d > 0; --d) {
            stride_[d - 1] = run;
            run *= size_[d - 1];
        }
    }

    template <typename... Indices>
    constexpr GKO_ACC_ATTRIBUTES GKO_ACC_INLINE value_type& operator()(
        Indices... indices) const
    {
        static_assert(sizeof...(Indices) == Dimensionality,
                      "one index per dimension");
        const size_type idx[Dimensionality] = {
            static_cast<size_type>(indices)...};
        size_type offset = idx[Dimensionality - 2] +
                           idx[Dimensionality - 1] * stride_[Dimensionality - 2];
        for (std::size_t d = 0; d + 2 < Dimensionality; ++d) {
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
