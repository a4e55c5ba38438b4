// elb200 host layer, part 4: the reference's BINARY / BINARY_FLAT matrix files to and from device-resident
// DistMatrix objects (src/io/Read/BinaryFlat.hpp:37-102, Read/Binary.hpp:44-..., Write/Binary.hpp:16-36,
// Write/BinaryFlat.hpp:16-33, Write.cpp:46-63).  Same bytes on disk as the reference; see csrc/host/io.cpp.
#pragma once
#include <string>

#include "elb200/core.hpp"

namespace El {

enum FileFormat { BINARY = 3, BINARY_FLAT = 4 };  // values of include/El/core/types.hpp:494-510 (AUTO, ASCII, ASCII_MATLAB, BINARY, BINARY_FLAT, ...)

namespace read {
// column-major entries of T, no header: the caller states the dimensions
template <typename T>
void BinaryFlat(AbstractDistMatrix<T>& A, Int height, Int width, const std::string& filename);
// two El::Int header words (height, width), then the entries
template <typename T>
void Binary(AbstractDistMatrix<T>& A, const std::string& filename);
}  // namespace read

namespace write {
template <typename T>
void Binary(const AbstractDistMatrix<T>& A, const std::string& basename = "matrix");       // basename + ".bin"
template <typename T>
void BinaryFlat(const AbstractDistMatrix<T>& A, const std::string& basename = "matrix");   // basename + ".dat"
}  // namespace write

}  // namespace El
