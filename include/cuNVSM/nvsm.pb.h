// cuNVSM/nvsm.pb.h — plain-struct stand-ins for the protobuf messages of proto/nvsm.proto
// (protoc is not part of this build). Accessor names and defaults follow the generated code
// the reference's callers use: desc.word_repr_size(), desc.transform_desc().nonlinearity(),
// train_config.update_method().type(), .adam_conf().mode(), mutable_*(), set_*().
#ifndef CUNVSM_B200_NVSM_PB_H
#define CUNVSM_B200_NVSM_PB_H

#include <cstdint>
#include <istream>
#include <iterator>
#include <ostream>
#include <string>
#include <vector>

namespace lse {

class ModelDesc {
 public:
  class TransformDesc {
   public:
    enum Nonlinearity { TANH = 0, HARD_TANH = 1 };
    bool batch_normalization() const { return batch_normalization_; }
    void set_batch_normalization(bool v) { batch_normalization_ = v; }
    Nonlinearity nonlinearity() const { return nonlinearity_; }
    void set_nonlinearity(Nonlinearity v) { nonlinearity_ = v; }
   private:
    bool batch_normalization_ = false;
    Nonlinearity nonlinearity_ = TANH;
  };

  int word_repr_size() const { return word_repr_size_; }
  void set_word_repr_size(int v) { word_repr_size_ = v; }
  int entity_repr_size() const { return entity_repr_size_; }
  void set_entity_repr_size(int v) { entity_repr_size_ = v; }
  const TransformDesc& transform_desc() const { return transform_desc_; }
  TransformDesc* mutable_transform_desc() { return &transform_desc_; }
  bool clip_sigmoid() const { return clip_sigmoid_; }
  void set_clip_sigmoid(bool v) { clip_sigmoid_ = v; }
  bool bias_negative_samples() const { return bias_negative_samples_; }
  void set_bias_negative_samples(bool v) { bias_negative_samples_ = v; }
  bool l2_normalize_phrase_reprs() const { return l2_normalize_phrase_reprs_; }
  void set_l2_normalize_phrase_reprs(bool v) { l2_normalize_phrase_reprs_ = v; }
  bool l2_normalize_entity_reprs() const { return l2_normalize_entity_reprs_; }
  void set_l2_normalize_entity_reprs(bool v) { l2_normalize_entity_reprs_ = v; }

 private:
  int word_repr_size_ = 0, entity_repr_size_ = 0;
  TransformDesc transform_desc_;
  bool clip_sigmoid_ = false, bias_negative_samples_ = false;
  bool l2_normalize_phrase_reprs_ = false, l2_normalize_entity_reprs_ = false;
};

class TrainConfig {
 public:
  enum UpdateMethod { SGD = 0, ADAGRAD = 1, ADAM = 2 };

  class UpdateMethodConf {
   public:
    class AdamConf {
     public:
      enum AdamMode { NONE = 0, SPARSE = 1, DENSE_UPDATE = 2, DENSE_UPDATE_DENSE_VARIANCE = 3 };
      AdamMode mode() const { return mode_; }
      void set_mode(AdamMode m) { mode_ = m; }
     private:
      AdamMode mode_ = NONE;
    };
    UpdateMethod type() const { return type_; }
    void set_type(UpdateMethod t) { type_ = t; }
    const AdamConf& adam_conf() const { return adam_conf_; }
    AdamConf* mutable_adam_conf() { return &adam_conf_; }
    void CopyFrom(const UpdateMethodConf& o) { *this = o; }
   private:
    UpdateMethod type_ = SGD;
    AdamConf adam_conf_;
  };

  int num_epochs() const { return num_epochs_; }
  void set_num_epochs(int v) { num_epochs_ = v; }
  int batch_size() const { return batch_size_; }
  void set_batch_size(int v) { batch_size_ = v; }
  int window_size() const { return window_size_; }
  void set_window_size(int v) { window_size_ = v; }
  int num_random_entities() const { return num_random_entities_; }
  void set_num_random_entities(int v) { num_random_entities_ = v; }
  float regularization_lambda() const { return regularization_lambda_; }
  void set_regularization_lambda(float v) { regularization_lambda_ = v; }
  float learning_rate() const { return learning_rate_; }
  void set_learning_rate(float v) { learning_rate_ = v; }
  const UpdateMethodConf& update_method() const { return update_method_; }
  UpdateMethodConf* mutable_update_method() { return &update_method_; }
  bool no_shuffle() const { return no_shuffle_; }
  void set_no_shuffle(bool v) { no_shuffle_ = v; }
  float text_entity_weight() const { return text_entity_weight_; }
  void set_text_entity_weight(float v) { text_entity_weight_ = v; }
  float entity_entity_weight() const { return entity_entity_weight_; }
  void set_entity_entity_weight(float v) { entity_entity_weight_ = v; }
  float term_term_weight() const { return term_term_weight_; }
  void set_term_term_weight(float v) { term_term_weight_ = v; }

 private:
  int num_epochs_ = 0, batch_size_ = 0, window_size_ = 0, num_random_entities_ = 0;
  float regularization_lambda_ = 0.f, learning_rate_ = 0.f;
  UpdateMethodConf update_method_;
  bool no_shuffle_ = false;
  float text_entity_weight_ = 1.f, entity_entity_weight_ = 0.f, term_term_weight_ = 0.f;
};

// Model serialization (proto/nvsm.proto:91-108). SerializeToOstream / ParseFromIstream speak the protobuf wire format
// (proto3: varint scalars, zero values omitted, sub-messages length-delimited), so `<output>_meta` files are readable by
// the reference's py/nvsm/base.py:load_meta (Metadata.ParseFromString) and the reference's own files parse here.
class Metadata {
 public:
  class TermInfo {
   public:
    int index_term_id() const { return index_term_id_; }
    void set_index_term_id(int v) { index_term_id_ = v; }
    int model_term_id() const { return model_term_id_; }
    void set_model_term_id(int v) { model_term_id_ = v; }
    int term_frequency() const { return term_frequency_; }
    void set_term_frequency(int v) { term_frequency_ = v; }
   private:
    int index_term_id_ = 0, model_term_id_ = 0, term_frequency_ = 0;
  };
  class ObjectInfo {
   public:
    int index_object_id() const { return index_object_id_; }
    void set_index_object_id(int v) { index_object_id_ = v; }
    int model_object_id() const { return model_object_id_; }
    void set_model_object_id(int v) { model_object_id_ = v; }
   private:
    int index_object_id_ = 0, model_object_id_ = 0;
  };

  TermInfo* add_term() { term_.emplace_back(); return &term_.back(); }
  int term_size() const { return static_cast<int>(term_.size()); }
  const TermInfo& term(int i) const { return term_[i]; }
  ObjectInfo* add_object() { object_.emplace_back(); return &object_.back(); }
  int object_size() const { return static_cast<int>(object_.size()); }
  const ObjectInfo& object(int i) const { return object_[i]; }
  int total_terms() const { return total_terms_; }
  void set_total_terms(int v) { total_terms_ = v; }

  std::string SerializeAsString() const {
    std::string out;
    for (const TermInfo& t : term_) {
      std::string sub;
      put_int32(&sub, 1, t.index_term_id()); put_int32(&sub, 2, t.model_term_id()); put_int32(&sub, 3, t.term_frequency());
      put_varint(&out, (1u << 3) | 2u); put_varint(&out, sub.size()); out += sub;
    }
    for (const ObjectInfo& o : object_) {
      std::string sub;
      put_int32(&sub, 1, o.index_object_id()); put_int32(&sub, 2, o.model_object_id());
      put_varint(&out, (2u << 3) | 2u); put_varint(&out, sub.size()); out += sub;
    }
    put_int32(&out, 3, total_terms_);
    return out;
  }
  bool SerializeToOstream(std::ostream* const os) const {
    const std::string bytes = SerializeAsString();
    os->write(bytes.data(), static_cast<std::streamsize>(bytes.size()));
    return os->good();
  }
  bool ParseFromString(const std::string& bytes) {
    term_.clear(); object_.clear(); total_terms_ = 0;
    size_t pos = 0;
    while (pos < bytes.size()) {
      uint64_t key = 0;
      if (!get_varint(bytes, &pos, bytes.size(), &key)) return false;
      const unsigned field = static_cast<unsigned>(key >> 3), wire = static_cast<unsigned>(key & 7u);
      if (wire == 2 && (field == 1 || field == 2)) {
        uint64_t len = 0;
        if (!get_varint(bytes, &pos, bytes.size(), &len) || len > bytes.size() - pos) return false;
        const size_t end = pos + static_cast<size_t>(len);
        int v[4] = {0, 0, 0, 0};
        while (pos < end) {
          uint64_t k = 0, x = 0;
          if (!get_varint(bytes, &pos, end, &k)) return false;
          if ((k & 7u) == 0) { if (!get_varint(bytes, &pos, end, &x)) return false; if ((k >> 3) >= 1 && (k >> 3) <= 3) v[k >> 3] = static_cast<int>(x); }
          else if (!skip(bytes, &pos, end, static_cast<unsigned>(k & 7u))) return false;
        }
        if (field == 1) { TermInfo* t = add_term(); t->set_index_term_id(v[1]); t->set_model_term_id(v[2]); t->set_term_frequency(v[3]); }
        else { ObjectInfo* o = add_object(); o->set_index_object_id(v[1]); o->set_model_object_id(v[2]); }
      } else if (wire == 0) {
        uint64_t x = 0;
        if (!get_varint(bytes, &pos, bytes.size(), &x)) return false;
        if (field == 3) total_terms_ = static_cast<int>(x);
      } else if (!skip(bytes, &pos, bytes.size(), wire)) {
        return false;
      }
    }
    return true;
  }
  bool ParseFromIstream(std::istream* const is) {
    const std::string bytes((std::istreambuf_iterator<char>(*is)), std::istreambuf_iterator<char>());
    return ParseFromString(bytes);
  }

 private:
  static void put_varint(std::string* out, uint64_t v) {
    while (v >= 0x80) { out->push_back(static_cast<char>((v & 0x7f) | 0x80)); v >>= 7; }
    out->push_back(static_cast<char>(v));
  }
  // int32 fields: sign-extended to 64 bits on the wire; proto3 omits zeros
  static void put_int32(std::string* out, unsigned field, int v) {
    if (v == 0) return;
    put_varint(out, (field << 3) | 0u);
    put_varint(out, static_cast<uint64_t>(static_cast<int64_t>(v)));
  }
  static bool get_varint(const std::string& in, size_t* pos, size_t end, uint64_t* v) {
    *v = 0;
    for (int shift = 0; shift < 64 && *pos < end; shift += 7) {
      const unsigned char c = static_cast<unsigned char>(in[(*pos)++]);
      *v |= static_cast<uint64_t>(c & 0x7f) << shift;
      if (!(c & 0x80)) return true;
    }
    return false;
  }
  static bool skip(const std::string& in, size_t* pos, size_t end, unsigned wire) {
    uint64_t x = 0;
    switch (wire) {
      case 0: return get_varint(in, pos, end, &x);
      case 1: if (end - *pos < 8) return false; *pos += 8; return true;
      case 2: if (!get_varint(in, pos, end, &x) || x > end - *pos) return false; *pos += static_cast<size_t>(x); return true;
      case 5: if (end - *pos < 4) return false; *pos += 4; return true;
      default: return false;
    }
  }
  std::vector<TermInfo> term_;
  std::vector<ObjectInfo> object_;
  int total_terms_ = 0;
};

typedef TrainConfig::UpdateMethodConf TrainConfig_UpdateMethodConf;
typedef TrainConfig::UpdateMethodConf::AdamConf TrainConfig_UpdateMethodConf_AdamConf;

}  // namespace lse

typedef lse::TrainConfig::UpdateMethodConf UpdateMethodConf;
using lse::TrainConfig;
static const lse::TrainConfig::UpdateMethod SGD = lse::TrainConfig::SGD;
static const lse::TrainConfig::UpdateMethod ADAGRAD = lse::TrainConfig::ADAGRAD;
static const lse::TrainConfig::UpdateMethod ADAM = lse::TrainConfig::ADAM;

#endif  // CUNVSM_B200_NVSM_PB_H
