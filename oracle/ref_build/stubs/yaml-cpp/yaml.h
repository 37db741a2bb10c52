// Stand-in for <yaml-cpp/yaml.h>.  The reference's parameter structs have constructors from a YAML::Node
// (C++/include/common.hpp:43-175) and C++/include/yaml_eigen_converter.hpp specialises YAML::convert<>.  The harness fills
// the structs through their default constructors, so these only have to COMPILE; calling one aborts.
#ifndef FBUS_REF_STUB_YAML
#define FBUS_REF_STUB_YAML
#include <cstdlib>
#include <string>
namespace YAML {
template <class T> struct convert;
class Node {
public:
    Node operator[](const char*) const { return Node(); }
    Node operator[](const std::string&) const { return Node(); }
    Node operator[](int) const { return Node(); }
    bool IsSequence() const { return false; }
    template <class T> T as() const { std::abort(); }
};
}  // namespace YAML
#endif
