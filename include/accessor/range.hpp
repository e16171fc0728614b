// gko::acc::range<Accessor> -- clean-room compat header (see
// accessor_config.hpp for provenance).  A range is a thin, trivially copyable
// view that forwards indexing and length queries to the accessor it wraps, so
// it can be passed by value as a __global__ kernel argument the way the
// reference does (cuda/gemv_kernels.cuh:191-192).
#pragma once

#include <utility>

#include "accessor_config.hpp"

namespace gko {
namespace acc {

template <typename Accessor>
class range {
public:
    using accessor = Accessor;
    static constexpr size_type dimensionality = Accessor::dimensionality;

    // `range(size, pointer, stride)` -- every argument is handed to the
    // accessor constructor unchanged.
    template <typename... Params>
    constexpr GKO_ACC_ATTRIBUTES explicit range(Params&&... params)
        : acc_(std::forward<Params>(params)...)
    {}

    range(const range&) = default;
    range(range&&) = default;
    ~range() = default;

    // element access: value for const storage, assignable proxy otherwise
    template <typename... Index>
    constexpr GKO_ACC_ATTRIBUTES GKO_ACC_INLINE auto operator()(
        Index... idx) const -> decltype(std::declval<const accessor&>()(
        static_cast<size_type>(idx)...))
    {
        return acc_(static_cast<size_type>(idx)...);
    }

    constexpr GKO_ACC_ATTRIBUTES GKO_ACC_INLINE size_type
    length(size_type dim) const
    {
        return acc_.length(dim);
    }

    constexpr GKO_ACC_ATTRIBUTES const accessor* operator->() const
    {
        return &acc_;
    }

    constexpr GKO_ACC_ATTRIBUTES const accessor& get_accessor() const
    {
        return acc_;
    }

private:
    accessor acc_;
};

}  // namespace acc
}  // namespace gko
