// glu/Reduce.hpp — glu::Reduce for CUDA device buffers (reference: glu/Reduce.hpp:42-135).
//
//     glu::Reduce reduce(glu::DataType_Uint, glu::ReduceOperator_Sum);
//     reduce(d_buffer, count);            // result in element 0 of d_buffer
//
// Same constructor and call operator as the reference; the GLuint SSBO name becomes a device pointer.
// The call only enqueues work on the object's stream (default stream unless set_stream() was called),
// exactly like the reference's glDispatchCompute sequence, and the result lands in element 0.
// Deviation (documented in DESIGN.md): elements 1..count-1 are left untouched, the reference overwrites
// some of them with partial results.
#ifndef GLU_B200_REDUCE_HPP
#define GLU_B200_REDUCE_HPP

#include "data_types.hpp"
#include "device_utils.hpp"

namespace glu
{
    /// The operators that can be used for the reduction (glu/Reduce.hpp:42-48).
    enum ReduceOperator
    {
        ReduceOperator_Sum = GLU_REDUCE_OPERATOR_SUM,
        ReduceOperator_Mul = GLU_REDUCE_OPERATOR_MUL,
        ReduceOperator_Min = GLU_REDUCE_OPERATOR_MIN,
        ReduceOperator_Max = GLU_REDUCE_OPERATOR_MAX
    };

    class Reduce
    {
    private:
        const DataType m_data_type;
        const ReduceOperator m_operator;
        DeviceBuffer m_tmp; // per-CTA partials + ticket, sized once (does not depend on count)
        glu_stream_t m_stream = nullptr;

    public:
        explicit Reduce(DataType data_type, ReduceOperator operator_) : m_data_type(data_type), m_operator(operator_)
        {
            (void) data_type_size(m_data_type); // "Invalid data type: %d"
            GLU_CHECK_ARGUMENT(int(m_operator) >= ReduceOperator_Sum && int(m_operator) <= ReduceOperator_Max,
                               "Invalid reduction operator: %d", int(m_operator));
            m_tmp.resize(glu_reduce_tmp_bytes(1, int(m_data_type)));
        }

        ~Reduce() = default;

        void set_stream(glu_stream_t stream) { m_stream = stream; }

        void operator()(DevicePtr buffer, size_t count)
        {
            GLU_CHECK_ARGUMENT(buffer, "Invalid buffer");
            GLU_CHECK_ARGUMENT(count > 0, "Count must be greater than zero");
            GLU_CHECK_STATUS(
                glu_reduce(buffer, count, int(m_data_type), int(m_operator), m_tmp.handle(), m_tmp.size(), m_stream));
        }
    };
} // namespace glu

#endif // GLU_B200_REDUCE_HPP
