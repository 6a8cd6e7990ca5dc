// methods/copy_functions.hpp -- receiver-row and column copies (reference methods/copy_functions.hpp:31-133)
#ifndef lzb_copy_functions_hpp
#define lzb_copy_functions_hpp

#include "../objects/dense_matrix.hpp"

// vec[vec_ind + c] = mat(lc, c)
template <typename type_t>
void copy_row_to_vector(unsigned int lc, const unsigned int vec_ind, Dense_matrix<type_t> &mat, Vector<type_t> &vec)
{
    if (mat.memory_space() == MemorySpace::CUDA) {
        AssertCuda(lz_copy_row(lanczos_context(), lc, (int)mat.n_cols(), reinterpret_cast<const double *>(mat.data()), (int64_t)mat.n_rows(),
                               reinterpret_cast<double *>(vec.data()), vec_ind));
    } else {
        for (unsigned int i = 0; i < mat.n_cols(); ++i) vec(vec_ind + i) = mat(lc + i * mat.n_rows());
    }
}
// first n_cols columns of mat1 -> mat2
template <typename type_t>
void copy_columns_to_matrix(Dense_matrix<type_t> &mat1, Dense_matrix<type_t> &mat2, const unsigned int n_cols)
{
    if (mat1.memory_space() == MemorySpace::CUDA) lzb::dcopy(mat2.data(), mat1.data(), mat1.n_rows() * n_cols * sizeof(type_t), LZ_D2D);
    else for (std::size_t i = 0; i < mat1.n_rows() * n_cols; ++i) mat2(i) = mat1(i);
}
template <typename type_t>
void copy_column_to_vector(const Dense_matrix<type_t> &mat, Vector<type_t> &vec, const unsigned int col)
{
    if (mat.memory_space() == MemorySpace::CUDA) lzb::dcopy(vec.data(), &(mat.data()[col * mat.n_rows()]), vec.size() * sizeof(type_t), LZ_D2D);
    else for (std::size_t i = 0; i < vec.size(); ++i) vec(i) = mat(i + col * vec.size());
}
template <typename type_t>
void copy_vector_to_column(const Vector<type_t> &vec, Dense_matrix<type_t> &mat, const unsigned int col)
{
    if (vec.memory_space() == MemorySpace::CUDA) lzb::dcopy(&(mat.data()[col * mat.n_rows()]), vec.data(), mat.n_rows() * sizeof(type_t), LZ_D2D);
    else for (std::size_t i = 0; i < vec.size(); ++i) mat(i + col * vec.size()) = vec(i);
}
template <typename type_t>
void copy_vector_element(const Vector<type_t> &vec1, const unsigned int src, Vector<type_t> &vec2, const unsigned int dest)
{
    if (vec1.memory_space() == MemorySpace::CUDA) lzb::dcopy(&(vec2.data()[dest]), &(vec1.data()[src]), sizeof(type_t), LZ_D2D);
    else vec2(dest) = vec1(src);
}

#endif
