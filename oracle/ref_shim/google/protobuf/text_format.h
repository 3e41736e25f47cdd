#ifndef REF_SHIM_PROTOBUF_TEXT_FORMAT_H
#define REF_SHIM_PROTOBUF_TEXT_FORMAT_H
#include <string>
#include "google/protobuf/message.h"
namespace google {
namespace protobuf {
class TextFormat {
 public:
  static bool PrintToString(const Message& message, std::string* output) {
      *output = message.DebugString();
      return true;
  }
};
}  // namespace protobuf
}  // namespace google
#endif
