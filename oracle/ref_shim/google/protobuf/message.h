// Stand-in for protobuf's Message (protobuf C++ headers / protoc are not installed).
#ifndef REF_SHIM_PROTOBUF_MESSAGE_H
#define REF_SHIM_PROTOBUF_MESSAGE_H
#include <string>
namespace google {
namespace protobuf {
class Message {
 public:
  virtual ~Message() {}
  virtual std::string DebugString() const { return std::string(); }
};
}  // namespace protobuf
}  // namespace google
#endif
