// gko::acc::scaled_reduced_row_major<D, ArithmeticType, StorageType, ScalarMask>
// -- clean-room (see accessor_config.hpp for provenance; interface from memory
// of upstream Ginkgo, NOT verifiable offline: the reference never instantiates
// it, its README (README.md:19) only says "all other accessors work
// accordingly").
//
// A reduced_row_major whose stored values are additionally multiplied by a
// scalar on read and divided by it on write:
//     value(i0..iD-1) = ArithmeticType(storage[...]) * scalar[s(i0..iD-1)]
// so that a low-precision (even integer) storage type can cover the dynamic
// range of each row / column.  ScalarMask selects the dimensions the scalar
// depends on: bit (D - 1 - d) set <=> index d takes part (written like the
// index tuple: on a 2-D accessor 0b10 = one scalar per ROW, 0b01 = one per
// COLUMN, 0b11 = one per element, 0 = a single scalar).  The scalars are a
// dense row-major array over the participating dimensions.
#pragma once

#include <array>
#include <cstdint>

#include "accessor_config.hpp"
#include "range.hpp"

namespace gko {
namespace acc {
namespace reference_class {

// Proxy for non-const storage: converts and scales on read, unscales and
// converts (one rounding) on write.
template <typename ArithmeticType, typename StorageType>
class scaled_reduced_storage {
public:
    using arithmetic_type = ArithmeticType;
    using storage_type = StorageType;

    constexpr GKO_ACC_ATTRIBUTES scaled_reduced_storage(storage_type* ptr,
                                                        arithmetic_type scalar)
        : ptr_(ptr), scalar_(scalar)
    {}
    scaled_reduced_storage(const scaled_reduced_storage&) = default;

    constexpr GKO_ACC_ATTRIBUTES GKO_ACC_INLINE operator arithmetic_type() const
    {
        return detail::storage_cast<arithmetic_type, storage_type>::apply(
                   *ptr_) *
               scalar_;
    }
    constexpr GKO_ACC_ATTRIBUTES GKO_ACC_INLINE arithmetic_type operator=(
        arithmetic_type value) const
    {
        *ptr_ = detail::storage_cast<storage_type, arithmetic_type>::apply(
            value / scalar_);
        return value;
    }
    constexpr GKO_ACC_ATTRIBUTES GKO_ACC_INLINE arithmetic_type operator=(
        const scaled_reduced_storage& other) const
    {
        return *this = static_cast<arithmetic_type>(other);
    }

#define GKO_ACC_SCALED_COMPOUND(op_)                                          \
    constexpr GKO_ACC_ATTRIBUTES GKO_ACC_INLINE arithmetic_type               \
    operator op_##=(arithmetic_type rhs) const                                \
    {                                                                         \
        return *this = static_cast<arithmetic_type>(*this) op_ rhs;           \
    }
    GKO_ACC_SCALED_COMPOUND(+)
    GKO_ACC_SCALED_COMPOUND(-)
    GKO_ACC_SCALED_COMPOUND(*)
    GKO_ACC_SCALED_COMPOUND(/)
#undef GKO_ACC_SCALED_COMPOUND

    constexpr GKO_ACC_ATTRIBUTES GKO_ACC_INLINE arithmetic_type
    operator-() const
    {
        return -static_cast<arithmetic_type>(*this);
    }

#define GKO_ACC_SCALED_BINARY(op_)                                            \
    friend constexpr GKO_ACC_ATTRIBUTES GKO_ACC_INLINE arithmetic_type        \
    operator op_(const scaled_reduced_storage& a,                             \
                 const scaled_reduced_storage& b)                             \
    {                                                                         \
        return static_cast<arithmetic_type>(a)                                \
            op_ static_cast<arithmetic_type>(b);                              \
    }                                                                         \
    friend constexpr GKO_ACC_ATTRIBUTES GKO_ACC_INLINE arithmetic_type        \
    operator op_(const scaled_reduced_storage& a, arithmetic_type b)          \
    {                                                                         \
        return static_cast<arithmetic_type>(a) op_ b;                         \
    }                                                                         \
    friend constexpr GKO_ACC_ATTRIBUTES GKO_ACC_INLINE arithmetic_type        \
    operator op_(arithmetic_type a, const scaled_reduced_storage& b)          \
    {                                                                         \
        return a op_ static_cast<arithmetic_type>(b);                         \
    }
    GKO_ACC_SCALED_BINARY(+)
    GKO_ACC_SCALED_BINARY(-)
    GKO_ACC_SCALED_BINARY(*)
    GKO_ACC_SCALED_BINARY(/)
#undef GKO_ACC_SCALED_BINARY

private:
    storage_type* ptr_;
    arithmetic_type scalar_;
};

}  // namespace reference_class


template <std::size_t Dimensionality, typename ArithmeticType,
          typename StorageType, std::uint64_t ScalarMask>
class scaled_reduced_row_major {
public:
    using arithmetic_type = typename std::remove_cv<ArithmeticType>::type;
    using storage_type = StorageType;
    static constexpr size_type dimensionality =
        static_cast<size_type>(Dimensionality);
    static constexpr std::uint64_t scalar_mask = ScalarMask;
    static constexpr bool is_const = std::is_const<storage_type>::value;
    using scalar_type =
        typename std::conditional<is_const, const arithmetic_type,
                                  arithmetic_type>::type;
    using const_accessor =
        scaled_reduced_row_major<Dimensionality, arithmetic_type,
                                 const storage_type, ScalarMask>;
    using length_type = std::array<size_type, Dimensionality>;

private:
    using bare_storage = typename std::remove_const<storage_type>::type;
    using proxy_type =
        reference_class::scaled_reduced_storage<arithmetic_type, storage_type>;

public:
    using reference_type =
        typename std::conditional<is_const, arithmetic_type, proxy_type>::type;

    static_assert(Dimensionality >= 1, "at least one dimension is required");
    static_assert(Dimensionality >= 64 || (ScalarMask >> Dimensionality) == 0,
                  "the scalar mask has bits beyond the dimensionality");

    // (size, storage, storage stride, scalars): the scalars are dense
    // row-major over the dimensions the mask selects
    template <typename SizeArray, typename StrideArray>
    constexpr GKO_ACC_ATTRIBUTES scaled_reduced_row_major(
        const SizeArray& size, storage_type* storage, const StrideArray& stride,
        scalar_type* scalar)
        : size_{}, stride_{}, storage_(storage), scalar_(scalar)
    {
        for (std::size_t d = 0; d < Dimensionality; ++d) {
            size_[d] = static_cast<size_type>(size[d]);
        }
        for (std::size_t d = 0; d + 1 < Dimensionality; ++d) {
            stride_[d] = static_cast<size_type>(stride[d]);
        }
    }
    template <typename SizeArray>
    constexpr GKO_ACC_ATTRIBUTES scaled_reduced_row_major(
        const SizeArray& size, storage_type* storage, scalar_type* scalar)
        : size_{}, stride_{}, storage_(storage), scalar_(scalar)
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

    constexpr GKO_ACC_ATTRIBUTES GKO_ACC_INLINE size_type
    length(size_type dim) const
    {
        return dim < dimensionality ? size_[dim] : size_type{1};
    }

    template <typename... Index>
    constexpr GKO_ACC_ATTRIBUTES GKO_ACC_INLINE size_type
    compute_index(Index... idx) const
    {
        static_assert(sizeof...(Index) == Dimensionality,
                      "number of indices must match the dimensionality");
        const size_type ids[Dimensionality] = {static_cast<size_type>(idx)...};
        size_type lin = ids[Dimensionality - 1];
        for (std::size_t d = 0; d + 1 < Dimensionality; ++d) {
            lin += ids[d] * stride_[d];
        }
        return lin;
    }

    // position in the dense scalar array: row-major over the masked dimensions
    template <typename... Index>
    constexpr GKO_ACC_ATTRIBUTES GKO_ACC_INLINE size_type
    compute_scalar_index(Index... idx) const
    {
        const size_type ids[Dimensionality] = {static_cast<size_type>(idx)...};
        size_type lin = 0;
        for (std::size_t d = 0; d < Dimensionality; ++d) {
            if ((ScalarMask >> (Dimensionality - 1 - d)) & 1u) {
                lin = lin * size_[d] + ids[d];
            }
        }
        return lin;
    }

    template <typename... Index, bool C = is_const>
    constexpr GKO_ACC_ATTRIBUTES GKO_ACC_INLINE
        typename std::enable_if<C, arithmetic_type>::type
        operator()(Index... idx) const
    {
        return detail::storage_cast<arithmetic_type, bare_storage>::apply(
                   storage_[compute_index(idx...)]) *
               scalar_[compute_scalar_index(idx...)];
    }
    template <typename... Index, bool C = is_const>
    constexpr GKO_ACC_ATTRIBUTES GKO_ACC_INLINE
        typename std::enable_if<!C, proxy_type>::type
        operator()(Index... idx) const
    {
        return proxy_type{storage_ + compute_index(idx...),
                          scalar_[compute_scalar_index(idx...)]};
    }

    // writes one scalar (the values stored under it are NOT rescaled)
    template <typename... Index, bool C = is_const>
    constexpr GKO_ACC_ATTRIBUTES GKO_ACC_INLINE
        typename std::enable_if<!C, arithmetic_type>::type
        write_scalar_masked(arithmetic_type value, Index... idx) const
    {
        return scalar_[compute_scalar_index(idx...)] = value;
    }
    template <typename... Index>
    constexpr GKO_ACC_ATTRIBUTES GKO_ACC_INLINE arithmetic_type
    read_scalar_masked(Index... idx) const
    {
        return scalar_[compute_scalar_index(idx...)];
    }

    constexpr GKO_ACC_ATTRIBUTES const_accessor to_const() const
    {
        return const_accessor{size_, storage_, stride_, scalar_};
    }
    constexpr GKO_ACC_ATTRIBUTES size_type get_stride(size_type d) const
    {
        return stride_[d];
    }
    constexpr GKO_ACC_ATTRIBUTES storage_type* get_stored_data() const
    {
        return storage_;
    }
    constexpr GKO_ACC_ATTRIBUTES const bare_storage* get_const_storage() const
    {
        return storage_;
    }
    constexpr GKO_ACC_ATTRIBUTES scalar_type* get_scalar() const
    {
        return scalar_;
    }
    constexpr GKO_ACC_ATTRIBUTES const arithmetic_type* get_const_scalar() const
    {
        return scalar_;
    }

private:
    size_type size_[Dimensionality];
    size_type stride_[Dimensionality > 1 ? Dimensionality - 1 : 1];
    storage_type* storage_;
    scalar_type* scalar_;
};

}  // namespace acc
}  // namespace gko
