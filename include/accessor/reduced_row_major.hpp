// gko::acc::reduced_row_major<D, ArithmeticType, StorageType> -- clean-room
// compat header (see accessor_config.hpp for provenance).
//
// A row-major view whose elements are *stored* as StorageType and *computed
// with* as ArithmeticType.  Reading converts storage -> arithmetic; writing
// through the proxy converts arithmetic -> storage with one rounding step.
// Surface required by the reference:
//   - ::arithmetic_type, ::storage_type, ::const_accessor, ::dimensionality
//   - ctor (std::array<size_type, D> size, StorageType* data,
//           std::array<size_type, D-1> stride)
//   - operator()(i0, ..., iD-1) const   -> arithmetic_type (const storage)
//                                          or an assignable proxy
//   - length(dim), get_stride(), get_size(), get_stored_data(),
//     get_const_storage()
#pragma once

#include <array>

#include "accessor_config.hpp"
#include "range.hpp"

namespace gko {
namespace acc {
namespace reference_class {

// Proxy returned for non-const storage.  Converts on read and on write, never
// holds a value itself.  All arithmetic on it happens in ArithmeticType.
template <typename ArithmeticType, typename StorageType>
class reduced_storage {
public:
    using arithmetic_type = ArithmeticType;
    using storage_type = StorageType;

    reduced_storage() = delete;
    constexpr GKO_ACC_ATTRIBUTES explicit reduced_storage(storage_type* ptr)
        : ptr_(ptr)
    {}
    reduced_storage(const reduced_storage&) = default;
    reduced_storage(reduced_storage&&) = default;
    ~reduced_storage() = default;

    constexpr GKO_ACC_ATTRIBUTES GKO_ACC_INLINE operator arithmetic_type() const
    {
        return detail::storage_cast<arithmetic_type, storage_type>::apply(
            *ptr_);
    }

    constexpr GKO_ACC_ATTRIBUTES GKO_ACC_INLINE arithmetic_type operator=(
        arithmetic_type value) const
    {
        *ptr_ =
            detail::storage_cast<storage_type, arithmetic_type>::apply(value);
        return value;
    }

    // proxy = proxy copies the *value*, not the pointer
    constexpr GKO_ACC_ATTRIBUTES GKO_ACC_INLINE arithmetic_type operator=(
        const reduced_storage& other) const
    {
        return *this = static_cast<arithmetic_type>(other);
    }
    constexpr GKO_ACC_ATTRIBUTES GKO_ACC_INLINE arithmetic_type operator=(
        reduced_storage&& other) const
    {
        return *this = static_cast<arithmetic_type>(other);
    }

#define GKO_ACC_PROXY_COMPOUND(op_)                                         \
    constexpr GKO_ACC_ATTRIBUTES GKO_ACC_INLINE arithmetic_type operator op_##=( \
        arithmetic_type rhs) const                                          \
    {                                                                       \
        return *this = static_cast<arithmetic_type>(*this) op_ rhs;         \
    }
    GKO_ACC_PROXY_COMPOUND(+)
    GKO_ACC_PROXY_COMPOUND(-)
    GKO_ACC_PROXY_COMPOUND(*)
    GKO_ACC_PROXY_COMPOUND(/)
#undef GKO_ACC_PROXY_COMPOUND

    constexpr GKO_ACC_ATTRIBUTES GKO_ACC_INLINE arithmetic_type
    operator-() const
    {
        return -static_cast<arithmetic_type>(*this);
    }

#define GKO_ACC_PROXY_BINARY(op_)                                           \
    friend constexpr GKO_ACC_ATTRIBUTES GKO_ACC_INLINE arithmetic_type      \
    operator op_(const reduced_storage& a, const reduced_storage& b)        \
    {                                                                       \
        return static_cast<arithmetic_type>(a)                              \
            op_ static_cast<arithmetic_type>(b);                            \
    }                                                                       \
    friend constexpr GKO_ACC_ATTRIBUTES GKO_ACC_INLINE arithmetic_type      \
    operator op_(const reduced_storage& a, arithmetic_type b)               \
    {                                                                       \
        return static_cast<arithmetic_type>(a) op_ b;                       \
    }                                                                       \
    friend constexpr GKO_ACC_ATTRIBUTES GKO_ACC_INLINE arithmetic_type      \
    operator op_(arithmetic_type a, const reduced_storage& b)               \
    {                                                                       \
        return a op_ static_cast<arithmetic_type>(b);                       \
    }
    GKO_ACC_PROXY_BINARY(+)
    GKO_ACC_PROXY_BINARY(-)
    GKO_ACC_PROXY_BINARY(*)
    GKO_ACC_PROXY_BINARY(/)
#undef GKO_ACC_PROXY_BINARY

private:
    storage_type* ptr_;
};

}  // namespace reference_class


template <std::size_t Dimensionality, typename ArithmeticType,
          typename StorageType>
class reduced_row_major {
public:
    using arithmetic_type = typename std::remove_cv<ArithmeticType>::type;
    using storage_type = StorageType;
    static constexpr size_type dimensionality =
        static_cast<size_type>(Dimensionality);
    static constexpr bool is_const = std::is_const<storage_type>::value;
    using const_accessor =
        reduced_row_major<Dimensionality, arithmetic_type, const storage_type>;
    using length_type = std::array<size_type, Dimensionality>;
    using storage_stride_type =
        std::array<size_type, (Dimensionality > 0 ? Dimensionality - 1 : 0)>;

private:
    using bare_storage = typename std::remove_const<storage_type>::type;
    using proxy_type =
        reference_class::reduced_storage<arithmetic_type, storage_type>;

public:
    using reference_type =
        typename std::conditional<is_const, arithmetic_type, proxy_type>::type;

    static_assert(Dimensionality >= 1, "at least one dimension is required");

    // (size, storage, stride) -- the signature the reference uses.  Any array
    // type with operator[] is accepted for size/stride (the reference passes a
    // `const std::array<std::int64_t, 2>` and a `std::array<size_type, 1>`).
    template <typename SizeArray, typename StrideArray>
    constexpr GKO_ACC_ATTRIBUTES reduced_row_major(const SizeArray& size,
                                                   storage_type* storage,
                                                   const StrideArray& stride)
        : size_{}, stride_{}, storage_(storage)
    {
        for (std::size_t d = 0; d < Dimensionality; ++d) {
            size_[d] = static_cast<size_type>(size[d]);
        }
        for (std::size_t d = 0; d + 1 < Dimensionality; ++d) {
            stride_[d] = static_cast<size_type>(stride[d]);
        }
    }

    // (size, storage): contiguous rows
    template <typename SizeArray>
    constexpr GKO_ACC_ATTRIBUTES reduced_row_major(const SizeArray& size,
                                                   storage_type* storage)
        : size_{}, stride_{}, storage_(storage)
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

    constexpr GKO_ACC_ATTRIBUTES reduced_row_major()
        : size_{}, stride_{}, storage_(nullptr)
    {}

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

    // const storage: arithmetic value
    template <typename... Index, bool C = is_const>
    constexpr GKO_ACC_ATTRIBUTES GKO_ACC_INLINE
        typename std::enable_if<C, arithmetic_type>::type
        operator()(Index... idx) const
    {
        return detail::storage_cast<arithmetic_type, bare_storage>::apply(
            storage_[compute_index(idx...)]);
    }

    // mutable storage: converting proxy
    template <typename... Index, bool C = is_const>
    constexpr GKO_ACC_ATTRIBUTES GKO_ACC_INLINE
        typename std::enable_if<!C, proxy_type>::type
        operator()(Index... idx) const
    {
        return proxy_type{storage_ + compute_index(idx...)};
    }

    constexpr GKO_ACC_ATTRIBUTES const_accessor to_const() const
    {
        return const_accessor{size_, storage_, stride_};
    }

    constexpr GKO_ACC_ATTRIBUTES const size_type* get_size_data() const
    {
        return size_;
    }
    GKO_ACC_ATTRIBUTES length_type get_size() const
    {
        length_type out{};
        for (std::size_t d = 0; d < Dimensionality; ++d) {
            out[d] = size_[d];
        }
        return out;
    }
    GKO_ACC_ATTRIBUTES storage_stride_type get_stride() const
    {
        storage_stride_type out{};
        for (std::size_t d = 0; d + 1 < Dimensionality; ++d) {
            out[d] = stride_[d];
        }
        return out;
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

private:
    size_type size_[Dimensionality];
    size_type stride_[Dimensionality > 1 ? Dimensionality - 1 : 1];
    storage_type* storage_;
};

}  // namespace acc
}  // namespace gko
