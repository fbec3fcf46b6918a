// glu/BlellochScan.hpp — glu::BlellochScan for CUDA device buffers (reference: glu/BlellochScan.hpp:79-190).
//
//     glu::BlellochScan scan(glu::DataType_Uint);
//     scan(d_buffer, count);                      // in-place exclusive prefix sum
//     scan(d_buffer, count, num_partitions);      // ... of num_partitions adjacent segments of `count`
//
// The name is kept for source compatibility; the algorithm underneath is a single-pass chained scan
// with decoupled look-back (csrc/glu_scan.cu), not the Blelloch tree.  On power-of-two counts (the only
// ones the reference accepts, glu/BlellochScan.hpp:134) the results are identical; other counts work too.
#ifndef GLU_B200_BLELLOCHSCAN_HPP
#define GLU_B200_BLELLOCHSCAN_HPP

#include "Reduce.hpp" // the reference's BlellochScan.hpp pulls Reduce.hpp in as well (glu/BlellochScan.hpp:6)
#include "data_types.hpp"

namespace glu
{
    class BlellochScan
    {
    private:
        const DataType m_data_type;
        DeviceBuffer m_tmp; // look-back state, grow-only
        glu_stream_t m_stream = nullptr;

    public:
        explicit BlellochScan(DataType data_type) : m_data_type(data_type) { (void) data_type_size(m_data_type); }

        ~BlellochScan() = default;

        void set_stream(glu_stream_t stream) { m_stream = stream; }

        /// Grow-only pre-sizing of the look-back state (optional; operator() does it on demand).
        void prepare_internal_buffers(size_t count, size_t num_partitions = 1)
        {
            const size_t need = glu_scan_exclusive_tmp_bytes(count, num_partitions, int(m_data_type));
            if (m_tmp.size() < need)
            {
                m_tmp.resize(need, false);
#ifdef GLU_VERBOSE
                std::printf("[BlellochScan] Look-back state reallocated to: %zu\n", need);
#endif
            }
        }

        void operator()(DevicePtr buffer, size_t count, size_t num_partitions = 1)
        {
            GLU_CHECK_ARGUMENT(buffer, "Invalid buffer");
            GLU_CHECK_ARGUMENT(count > 0, "Count must be greater than zero");
            GLU_CHECK_ARGUMENT(num_partitions >= 1, "Num of partitions must be >= 1");
            prepare_internal_buffers(count, num_partitions);
            GLU_CHECK_STATUS(glu_scan_exclusive(buffer, count, num_partitions, int(m_data_type), m_tmp.handle(),
                                                m_tmp.size(), m_stream));
        }
    };
} // namespace glu

#endif // GLU_B200_BLELLOCHSCAN_HPP
