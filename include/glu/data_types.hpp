// glu/data_types.hpp — the element types the primitives accept (reference: glu/data_types.hpp:8-22,
// same enumerator names and values).  The reference turns the enum into a GLSL type name for its
// runtime shader specialisation (to_glsl_type_str, :24-44); here the kernels are compile-time template
// instantiations selected by the enum inside libglu_b200.so, so what callers need instead is the
// element size (std430 array stride) and a printable name.
#ifndef GLU_B200_DATA_TYPES_HPP
#define GLU_B200_DATA_TYPES_HPP

#include <cstddef>

#include "errors.hpp"

namespace glu
{
    enum DataType
    {
        DataType_Float = GLU_DATA_TYPE_FLOAT,
        DataType_Double = GLU_DATA_TYPE_DOUBLE,
        DataType_Int = GLU_DATA_TYPE_INT,
        DataType_Uint = GLU_DATA_TYPE_UINT,
        DataType_Vec2 = GLU_DATA_TYPE_VEC2,
        DataType_Vec4 = GLU_DATA_TYPE_VEC4,
        DataType_DVec2 = GLU_DATA_TYPE_DVEC2,
        DataType_DVec4 = GLU_DATA_TYPE_DVEC4,
        DataType_UVec2 = GLU_DATA_TYPE_UVEC2,
        DataType_UVec4 = GLU_DATA_TYPE_UVEC4,
        DataType_IVec2 = GLU_DATA_TYPE_IVEC2,
        DataType_IVec4 = GLU_DATA_TYPE_IVEC4
    };

    /// Same strings the reference's to_glsl_type_str returns; an invalid id is fatal ("Invalid data type: %d").
    inline const char* to_glsl_type_str(DataType data_type)
    {
        static const char* const k_names[] = {"float", "double", "int",   "uint",  "vec2",  "vec4",
                                              "dvec2", "dvec4",  "uvec2", "uvec4", "ivec2", "ivec4"};
        GLU_CHECK_ARGUMENT(int(data_type) >= 0 && int(data_type) < 12, "Invalid data type: %d", int(data_type));
        return k_names[int(data_type)];
    }

    /// Bytes per element (4 / 8 / 16 / 32).
    inline size_t data_type_size(DataType data_type)
    {
        const size_t size = glu_data_type_size(int(data_type));
        GLU_CHECK_ARGUMENT(size != 0, "Invalid data type: %d", int(data_type));
        return size;
    }
} // namespace glu

#endif // GLU_B200_DATA_TYPES_HPP
