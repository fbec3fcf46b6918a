// glu/device_utils.hpp — what survives of the reference's glu/gl_utils.hpp once the OpenGL plumbing is
// gone:
//   * DeviceBuffer: the ShaderStorageBuffer role (glu/gl_utils.hpp:146-246) for CUDA device memory —
//     move-only RAII, construct from host data, resize / clear / write_data / get_data<T>, handle();
//   * measure_elapsed_time: the measure_gl_elapsed_time role (glu/gl_utils.hpp:249-265) with CUDA events;
//   * the integer helpers div_ceil / is_power_of_2 / next_power_of_2 / log32_* (glu/gl_utils.hpp:267-302)
//     with the reference's edge-case behaviour (is_power_of_2(0) is true, next_power_of_2(0) is 0);
//   * print_buffer / print_buffer_hex / print_stl_container debug dumps (glu/gl_utils.hpp:304-329).
// Shader / Program (runtime GLSL compilation) have no counterpart: kernels are compiled ahead of time.
// Everything goes through the C ABI, so this header needs no CUDA toolkit to compile.
#ifndef GLU_B200_DEVICE_UTILS_HPP
#define GLU_B200_DEVICE_UTILS_HPP

#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <functional>
#include <string>
#include <utility>
#include <vector>

#include "errors.hpp"

namespace glu
{
    /// The "buffer handle" of this build: a CUDA device pointer (the reference passes GLuint SSBO names).
    using DevicePtr = void*;

    inline void copy_buffer(DevicePtr src, DevicePtr dst, size_t size, size_t src_offset = 0, size_t dst_offset = 0,
                            glu_stream_t stream = nullptr)
    {
        GLU_CHECK_STATUS(glu_memcpy_d2d(static_cast<char*>(dst) + dst_offset, static_cast<const char*>(src) + src_offset,
                                        size, stream));
    }

    /// Move-only owner of one device allocation.
    class DeviceBuffer
    {
    private:
        DevicePtr m_handle = nullptr;
        size_t m_size = 0;

        void release()
        {
            if (m_handle)
                glu_free(m_handle);
            m_handle = nullptr;
            m_size = 0;
        }

    public:
        explicit DeviceBuffer(size_t initial_size = 0)
        {
            if (initial_size > 0)
                resize(initial_size, false);
        }

        explicit DeviceBuffer(const void* data, size_t size)
        {
            GLU_CHECK_ARGUMENT(data, "DeviceBuffer: null host data");
            GLU_CHECK_ARGUMENT(size > 0, "DeviceBuffer: size must be greater than zero");
            resize(size, false);
            write_data(data, size);
        }

        template<typename T>
        explicit DeviceBuffer(const std::vector<T>& data) : DeviceBuffer(data.data(), data.size() * sizeof(T))
        {
        }

        DeviceBuffer(const DeviceBuffer&) = delete;
        DeviceBuffer& operator=(const DeviceBuffer&) = delete;
        DeviceBuffer(DeviceBuffer&& other) noexcept : m_handle(other.m_handle), m_size(other.m_size)
        {
            other.m_handle = nullptr;
            other.m_size = 0;
        }
        DeviceBuffer& operator=(DeviceBuffer&& other) noexcept
        {
            if (this != &other)
            {
                release();
                std::swap(m_handle, other.m_handle);
                std::swap(m_size, other.m_size);
            }
            return *this;
        }

        ~DeviceBuffer() { release(); }

        [[nodiscard]] DevicePtr handle() const { return m_handle; }
        [[nodiscard]] size_t size() const { return m_size; }
        template<typename T> [[nodiscard]] T* as() const { return static_cast<T*>(m_handle); }

        /// Grows or shrinks the allocation; with keep_data the common prefix survives.
        void resize(size_t size, bool keep_data = false)
        {
            if (size == m_size)
                return;
            DevicePtr fresh = nullptr;
            if (size > 0)
                GLU_CHECK_STATUS(glu_malloc(&fresh, size));
            if (keep_data && m_handle && fresh)
            {
                copy_buffer(m_handle, fresh, std::min(m_size, size));
                GLU_CHECK_STATUS(glu_stream_synchronize(nullptr));
            }
            if (m_handle)
                glu_free(m_handle);
            m_handle = fresh;
            m_size = size;
        }

        /// Fills the whole buffer with a repeated 32-bit value.
        void clear(uint32_t value)
        {
            if (m_size >= sizeof(uint32_t))
                GLU_CHECK_STATUS(glu_memset_u32(m_handle, value, m_size / sizeof(uint32_t), nullptr));
        }

        void write_data(const void* data, size_t size)
        {
            GLU_CHECK_ARGUMENT(size <= m_size, "DeviceBuffer::write_data: %zu bytes do not fit in %zu", size, m_size);
            GLU_CHECK_STATUS(glu_memcpy_h2d(m_handle, data, size, nullptr));
            GLU_CHECK_STATUS(glu_stream_synchronize(nullptr));
        }

        /// Downloads the whole buffer.  Uploads, downloads and clears run on the legacy default stream, which is
        /// ordered with every stream made by glu_stream_create (blocking streams) — the stream `set_stream` takes.
        /// Work enqueued on a cudaStreamNonBlocking stream of the caller's own must be synchronised by the caller.
        template<typename T> std::vector<T> get_data() const
        {
            GLU_CHECK_ARGUMENT(m_size % sizeof(T) == 0, "Size %zu isn't a multiple of %zu", m_size, sizeof(T));
            std::vector<T> result(m_size / sizeof(T));
            if (m_size)
            {
                GLU_CHECK_STATUS(glu_memcpy_d2h(result.data(), m_handle, m_size, nullptr));
                GLU_CHECK_STATUS(glu_stream_synchronize(nullptr));
            }
            return result;
        }
    };

    /// Name kept so that code written against the reference's buffer helper compiles unchanged.
    using ShaderStorageBuffer = DeviceBuffer;

    /// Device time, in nanoseconds, of whatever `callback` enqueues on `stream`.
    inline uint64_t measure_elapsed_time(const std::function<void()>& callback, glu_stream_t stream = nullptr)
    {
        glu_event_t begin = nullptr, end = nullptr;
        GLU_CHECK_STATUS(glu_event_create(&begin));
        GLU_CHECK_STATUS(glu_event_create(&end));
        GLU_CHECK_STATUS(glu_event_record(begin, stream));
        callback();
        GLU_CHECK_STATUS(glu_event_record(end, stream));
        GLU_CHECK_STATUS(glu_event_synchronize(end));
        float ms = 0.f;
        GLU_CHECK_STATUS(glu_event_elapsed_ms(&ms, begin, end));
        glu_event_destroy(begin);
        glu_event_destroy(end);
        return uint64_t(double(ms) * 1e6);
    }
    inline uint64_t measure_gl_elapsed_time(const std::function<void()>& callback) { return measure_elapsed_time(callback); }

    template<typename IntegerT> IntegerT log32_floor(IntegerT n) { return IntegerT(std::floor(std::log2(double(n)) / 5.0)); }
    template<typename IntegerT> IntegerT log32_ceil(IntegerT n) { return IntegerT(std::ceil(std::log2(double(n)) / 5.0)); }

    /// Exact integer ceil-division (the reference goes through double, which agrees for every size it can reach).
    template<typename IntegerT> IntegerT div_ceil(IntegerT n, IntegerT d) { return IntegerT(n / d + (n % d != 0 ? 1 : 0)); }

    /// True for powers of two and, like the reference, for 0.
    template<typename T> bool is_power_of_2(T n) { return (n & (n - 1)) == 0; }

    /// Smallest power of two >= n for 32-bit n; 0 -> 0 like the reference.
    template<typename IntegerT> IntegerT next_power_of_2(IntegerT n)
    {
        if (n == 0)
            return 0;
        IntegerT p = 1;
        while (p < n)
            p = IntegerT(p << 1);
        return p;
    }

    template<typename Iterator> void print_stl_container(Iterator begin, Iterator end)
    {
        for (size_t i = 0; begin != end; ++begin, ++i)
            std::printf("(%zu) %s, ", i, std::to_string(*begin).c_str());
        std::printf("\n");
    }

    template<typename T> void print_buffer(const DeviceBuffer& buffer)
    {
        const std::vector<T> data = buffer.get_data<T>();
        print_stl_container(data.begin(), data.end());
    }

    inline void print_buffer_hex(const DeviceBuffer& buffer)
    {
        const std::vector<uint32_t> data = buffer.get_data<uint32_t>();
        for (size_t i = 0; i < data.size(); i++)
            std::printf("(%zu) %08x, ", i, data[i]);
        std::printf("\n");
    }
} // namespace glu

#endif // GLU_B200_DEVICE_UTILS_HPP
