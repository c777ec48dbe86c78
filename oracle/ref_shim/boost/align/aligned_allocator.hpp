// Minimal stand-in for <boost/align/aligned_allocator.hpp>, written for the
// oracle build only (TEST INFRASTRUCTURE, not product code).  Boost is not
// installed in this image; the reference's SimdLib/allocator.hpp needs just
// boost::alignment::aligned_allocator<T, Alignment>.
#pragma once
#include <cstddef>
#include <cstdlib>
#include <new>

namespace boost { namespace alignment {

template <class T, std::size_t Alignment>
struct aligned_allocator
{
    using value_type = T;
    template <class U> struct rebind { using other = aligned_allocator<U, Alignment>; };

    aligned_allocator() noexcept = default;
    template <class U>
    aligned_allocator(const aligned_allocator<U, Alignment> &) noexcept {}

    T *allocate(std::size_t n)
    {
        constexpr std::size_t a = Alignment < sizeof(void *) ? sizeof(void *) : Alignment;
        std::size_t bytes = n * sizeof(T);
        bytes = (bytes + a - 1) / a * a;
        if (bytes == 0) bytes = a;
        void *p = nullptr;
        if (posix_memalign(&p, a, bytes) != 0) p = nullptr;
        if (!p) throw std::bad_alloc();
        return static_cast<T *>(p);
    }
    void deallocate(T *p, std::size_t) noexcept { std::free(p); }
};

template <class T, class U, std::size_t A>
bool operator==(const aligned_allocator<T, A> &, const aligned_allocator<U, A> &) noexcept { return true; }
template <class T, class U, std::size_t A>
bool operator!=(const aligned_allocator<T, A> &, const aligned_allocator<U, A> &) noexcept { return false; }

}} // namespace boost::alignment
