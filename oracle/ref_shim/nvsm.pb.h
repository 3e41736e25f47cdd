// Stand-in for the protoc-generated nvsm.pb.h (protoc / libprotobuf are not installed): plain classes with
// the accessor names the reference's sources call (proto/nvsm.proto), used only to compile the
// UNMODIFIED reference into oracle/_ref. Test infrastructure.
#ifndef REF_SHIM_NVSM_PB_H
#define REF_SHIM_NVSM_PB_H

#include <string>

#include "google/protobuf/message.h"

namespace lse {

class ModelDesc : public ::google::protobuf::Message {
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

class TrainConfig : public ::google::protobuf::Message {
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


class DataConfig : public ::google::protobuf::Message {};

class Metadata : public ::google::protobuf::Message {
 public:
  class TermInfo {};
  class ObjectInfo {};
};

typedef TrainConfig::UpdateMethodConf TrainConfig_UpdateMethodConf;
typedef TrainConfig::UpdateMethodConf::AdamConf TrainConfig_UpdateMethodConf_AdamConf;

}  // namespace lse

#endif  // REF_SHIM_NVSM_PB_H
