// The reference's Params / Storage / Updates / BatchNormalization class surface (include/cuNVSM/{params,storage,updates,
// cudnn_utils}.h of this repo, header-only over libnvsm_b200's nvsm_op_* / nvsm_updater_* entry points) driven the way the
// reference's own unit tests drive it, on the inputs those tests use:
//   cpp/updates_tests.cu:34-775   every GradientUpdater, the four (lambda, learning rate) parameterisations,
//                                 incl. the in-place gradient rewrites and the non-decaying Adam bias moments
//   cpp/model_tests.cu:52-339     get_average_representations (plain / weighted), Representations update (decay, scatter),
//                                 update_dense, Transform::transform;  :468-521 transform + batch-norm + tanh golden values
//   cpp/cudnn_utils_tests.cu:19-177  BatchNormalization forward (closed form), backward (golden grad_bias), in-place == out-of-place
// Expected values are recomputed here in double from the definitions (sums per object, window means, bias correction);
// literals are only used where the reference pins a literal. float32 library vs double expectation: 3e-5 relative
// (beta2 = 0.999 is 0.99900001 in float32, so every (1 - beta2) factor carries 1.3e-5, like the reference's float build).
// Prints one line per case and "CLASSES_TEST_OK" when everything held; exit status = number of failed cases.
#include <cmath>
#include <cstdio>
#include <functional>
#include <memory>
#include <string>
#include <vector>

#include "cuNVSM/model.h"
#include "cuNVSM/params.h"

typedef float FloatT;
typedef int32 IdxType;
typedef std::vector<double> Vec;

static int g_failed = 0, g_cases = 0;

static bool near(const std::vector<FloatT>& got, const Vec& want, const char* what, double rel = 3e-5, double abs_tol = 1e-6) {
  if (got.size() != want.size()) { std::printf("  %s: size %zu vs %zu\n", what, got.size(), want.size()); return false; }
  for (size_t i = 0; i < got.size(); ++i)
    if (!(std::fabs(got[i] - want[i]) <= abs_tol + rel * std::fabs(want[i]))) {
      std::printf("  %s[%zu]: got %.9g, expected %.9g\n", what, i, (double)got[i], want[i]);
      return false;
    }
  return true;
}

static void run_case(const std::string& name, const std::function<bool()>& body) {
  ++g_cases;
  const bool ok = body();
  if (!ok) ++g_failed;
  std::printf("%s %s\n", ok ? "ok  " : "FAIL", name.c_str());
}

// the reference's UpdatesTest fixture: storages start at a constant
template <typename T>
static std::unique_ptr<T> constant_storage(const FloatT value, const size_t first, const size_t second) {
  std::unique_ptr<T> s(new T(first, second, DefaultStream::get()));
  s->initialize_with_constant(value);
  return s;
}

// Inputs shared by the representation-updater cases of cpp/updates_tests.cu: two gradient columns of width 4.
static const Vec kG1 = {2.0, 2.5, 3.0, 4.0}, kG2 = {10.0, 11.0, 12.0, 13.0};
static double mean_sq(const Vec& g) { double s = 0; for (double x : g) s += x * x; return s / g.size(); }

// per-object sums of the scattered gradient: hits[o] = list of gradient columns landing on object o
static std::vector<Vec> scatter_sum(const size_t num_objects, const std::vector<std::pair<std::vector<long>, Vec>>& descs) {
  std::vector<Vec> out(num_objects, Vec(4, 0.0));
  for (const auto& d : descs)
    for (long id : d.first)
      for (int k = 0; k < 4; ++k) out[id][k] += d.second[k];
  return out;
}

static std::vector<FloatT> f(const Vec& v) { return std::vector<FloatT>(v.begin(), v.end()); }

int main() {
  Streams* const S = DefaultStream::get();
  const double lambdas[2] = {0.0, 0.1}, lrs[2] = {1.0, 0.5};   // INSTANTIATE_TEST_CASE_P(Regularization, ...) of the reference

  run_case("device_matrix: column-major round trip, fill, copy", [&] {
    device_matrix<FloatT> m(3, 2);
    bool ok = near(to_host(m), Vec(6, 0.0), "zero-initialised");
    to_device({1.f, 2.f, 3.f, 4.f, 5.f, 6.f}, &m);
    std::unique_ptr<device_matrix<FloatT>> c(m.copy());
    m.fillwith(nullptr, 7.0f);
    ok = ok && near(to_host(*c), {1, 2, 3, 4, 5, 6}, "copy") && near(to_host(m), Vec(6, 7.0), "fill");
    device_matrix<IdxType> idx(1, 3);
    to_device({9L, 0L, 1L}, &idx);
    const std::vector<IdxType> h = to_host(idx);
    return ok && h[0] == 9 && h[1] == 0 && h[2] == 1 && m.getRows() == 3 && m.getCols() == 2;
  });

  for (const double lam : lambdas)
    for (const double lr : lrs) {
      char tag[64];
      std::snprintf(tag, sizeof tag, " [lambda %.1f lr %.1f]", lam, lr);
      const std::string T(tag);
      Vec graw(24), gbias = {25.0, 26.0, 27.0};
      for (int i = 0; i < 24; ++i) graw[i] = i + 1.0;

      run_case("SGDTransformGradientUpdater" + T, [&] {
        auto st = constant_storage<TransformStorage<FloatT>>(5.0, 8, 3);
        SGDTransformGradientUpdater<FloatT> up;
        device_matrix<FloatT> gT(3, 8), gb(3, 1);
        to_device(f(graw), &gT); to_device(f(gbias), &gb);
        TransformStorage<FloatT>::GradientType desc = std::forward_as_tuple(gT, gb);
        up.update(st.get(), &desc, lr, lam, S);
        Vec wantT(24), wantb(3);
        for (int i = 0; i < 24; ++i) wantT[i] = 5.0 + lr * (graw[i] - lam * 5.0);
        for (int i = 0; i < 3; ++i) wantb[i] = 5.0 + lr * gbias[i];     // the bias is never regularised
        return near(to_host(*std::get<0>(st->get())), wantT, "transform") && near(to_host(*std::get<1>(st->get())), wantb, "bias");
      });

      run_case("SGDRepresentationsGradientUpdater, two descriptors" + T, [&] {
        auto st = constant_storage<RepresentationsStorage<FloatT, IdxType>>(5.0, 10, 4);
        SGDRepresentationsGradientUpdater<FloatT, IdxType> up;
        device_matrix<FloatT> g1(4, 1), g2(4, 1);
        to_device(f(kG1), &g1); to_device(f(kG2), &g2);
        device_matrix<IdxType> i1(1, 3), i2(1, 3);
        to_device({9L, 0L, 1L}, &i1); to_device({5L, 1L, 8L}, &i2);
        RepresentationsStorage<FloatT, IdxType>::GradientType desc = {std::forward_as_tuple(g1, i1, (size_t)3, nullptr),
                                                                      std::forward_as_tuple(g2, i2, (size_t)3, nullptr)};
        up.update(st.get(), &desc, lr, lam, S);
        const auto sums = scatter_sum(10, {{{9, 0, 1}, kG1}, {{5, 1, 8}, kG2}});
        Vec want;
        for (int o = 0; o < 10; ++o)
          for (int k = 0; k < 4; ++k) want.push_back(5.0 * (1.0 - lr * lam) + lr * sums[o][k]);
        return near(to_host(*st->get()), want, "representations");
      });

      run_case("AdagradTransformGradientUpdater: accumulators, in-place gradients" + T, [&] {
        const double eps = 1e-6;
        auto st = constant_storage<TransformStorage<FloatT>>(5.0, 8, 3);
        AdagradTransformGradientUpdater<FloatT> up(8, 3, S, eps);
        device_matrix<FloatT> gT(3, 8), gb(3, 1);
        to_device(f(graw), &gT); to_device(f(gbias), &gb);
        TransformStorage<FloatT>::GradientType desc = std::forward_as_tuple(gT, gb);
        up.update(st.get(), &desc, lr, lam, S);
        Vec acc(24), accb(3), gmod(24), gbmod(3), wantT(24), wantb(3);
        for (int i = 0; i < 24; ++i) { acc[i] = graw[i] * graw[i]; gmod[i] = graw[i] / std::sqrt(acc[i] + eps); wantT[i] = 5.0 * (1 - lam * lr) + lr * gmod[i]; }
        for (int i = 0; i < 3; ++i) { accb[i] = gbias[i] * gbias[i]; gbmod[i] = gbias[i] / std::sqrt(accb[i] + eps); wantb[i] = 5.0 + lr * gbmod[i]; }
        return near(up.state("acc"), acc, "acc") && near(up.state("acc_bias"), accb, "acc_bias") && near(to_host(gT), gmod, "grad_transform in place") &&
               near(to_host(gb), gbmod, "grad_bias in place") && near(to_host(*std::get<0>(st->get())), wantT, "transform") &&
               near(to_host(*std::get<1>(st->get())), wantb, "bias");
      });

      run_case("AdagradRepresentationsGradientUpdater: per-object scalars, window-mean divisor" + T, [&] {
        const double eps = 1e-6;
        auto st = constant_storage<RepresentationsStorage<FloatT, IdxType>>(5.0, 10, 4);
        AdagradRepresentationsGradientUpdater<FloatT, IdxType> up(10, S, eps);
        device_matrix<FloatT> g(4, 2);
        Vec both = kG1; both.insert(both.end(), kG2.begin(), kG2.end());
        to_device(f(both), &g);
        device_matrix<IdxType> idx(1, 6);
        const std::vector<long> ids = {9, 0, 1, 5, 1, 8};
        to_device(ids, &idx);
        RepresentationsStorage<FloatT, IdxType>::GradientType desc = {std::forward_as_tuple(g, idx, (size_t)3, nullptr)};
        up.update(st.get(), &desc, lr, lam, S);
        Vec acc(10, 0.0);
        const double ms[2] = {mean_sq(kG1), mean_sq(kG2)};   // 8.8125, 133.5
        for (int j = 0; j < 6; ++j) acc[ids[j]] += ms[j / 3];
        Vec gmod, table(40, 5.0 * (1 - lam * lr));
        for (int x = 0; x < 2; ++x) {
          double a = 0;
          for (int y = 0; y < 3; ++y) a += acc[ids[3 * x + y]];
          const double den = std::sqrt(a / 3.0 + eps);
          for (int k = 0; k < 4; ++k) {
            const double gk = (x ? kG2 : kG1)[k] / den;
            gmod.push_back(gk);
            for (int y = 0; y < 3; ++y) table[ids[3 * x + y] * 4 + k] += lr * gk;
          }
        }
        return near(up.state("acc"), acc, "acc") && near(to_host(g), gmod, "grad in place") && near(to_host(*st->get()), table, "representations");
      });

      run_case("AdamTransformGradientUpdater: two steps, in-place step direction, non-decaying bias moments" + T, [&] {
        const double eps = 1e-5, b1 = 0.9, b2 = 0.999;
        auto st = constant_storage<TransformStorage<FloatT>>(5.0, 8, 3);
        AdamTransformGradientUpdater<FloatT> up(8, 3, S, b1, b2, eps);
        Vec m(24, 0.0), v(24, 0.0), mb(3, 0.0), vb(3, 0.0), Tcur(24, 5.0);
        bool ok = true;
        for (int step = 1; step <= 2 && ok; ++step) {
          device_matrix<FloatT> gT(3, 8), gb(3, 1);
          to_device(f(graw), &gT); to_device(f(gbias), &gb);
          TransformStorage<FloatT>::GradientType desc = std::forward_as_tuple(gT, gb);
          up.update(st.get(), &desc, lr, lam, S);
          const double bc = std::sqrt(1.0 - std::pow(b2, step)) / (1.0 - std::pow(b1, step));
          Vec dir(24), dirb(3);
          for (int i = 0; i < 24; ++i) {
            const double g = graw[i] - lam * Tcur[i];
            m[i] = b1 * m[i] + (1 - b1) * g; v[i] = b2 * v[i] + (1 - b2) * g * g;
            dir[i] = bc * m[i] / (std::sqrt(v[i]) + eps);
            Tcur[i] += lr * dir[i];
          }
          for (int i = 0; i < 3; ++i) {   // no decay on the bias moments (cpp/updates_tests.cu:352-366,409-423)
            mb[i] += (1 - b1) * gbias[i]; vb[i] += (1 - b2) * gbias[i] * gbias[i];
            dirb[i] = bc * mb[i] / (std::sqrt(vb[i]) + eps);
          }
          ok = near(to_host(gT), dir, "grad_transform in place") && near(to_host(gb), dirb, "grad_bias in place") &&
               near(up.state("m_bias"), mb, "m_bias") && near(up.state("v_bias"), vb, "v_bias") && near(up.state("m"), m, "m") &&
               near(up.state("v"), v, "v") && near(to_host(*std::get<0>(st->get())), Tcur, "transform");
          // the literals the reference pins
          const Vec lit = step == 1 ? Vec{0.9999873510493572093, 0.99998783754154196846, 0.99998828799769046149}
                                    : Vec{1.0523589755648365962, 1.0523593375842164033, 1.0523596727875677015};
          const Vec litm = step == 1 ? Vec{2.5, 2.6, 2.7} : Vec{5.0, 5.2, 5.4};
          const Vec litv = step == 1 ? Vec{0.625, 0.676, 0.729} : Vec{1.25, 1.352, 1.458};
          ok = ok && near(to_host(gb), lit, "grad_bias literal") && near(up.state("m_bias"), litm, "m_bias literal") &&
               near(up.state("v_bias"), litv, "v_bias literal");
        }
        return ok;
      });

      // the three AdamRepresentationsGradientUpdater modes on the reference's 5-object table
      const std::vector<long> idsA = {4, 0, 1}, idsB = {3, 1, 2};
      auto adam_state = [&](const double b1, const double b2, std::vector<Vec>* sums, Vec* vscalar) {
        *sums = scatter_sum(5, {{idsA, kG1}, {idsB, kG2}});
        vscalar->assign(5, 0.0);
        for (long id : idsA) (*vscalar)[id] += (1 - b2) * mean_sq(kG1);
        for (long id : idsB) (*vscalar)[id] += (1 - b2) * mean_sq(kG2);
        (void)b1;
      };

      run_case("AdamRepresentationsGradientUpdater SPARSE: moments, window-averaged step in place" + T, [&] {
        const double eps = 1e-5, b1 = 0.9, b2 = 0.999;
        auto st = constant_storage<RepresentationsStorage<FloatT, IdxType>>(5.0, 5, 4);
        AdamConf conf; conf.set_mode(AdamConf::SPARSE);
        AdamRepresentationsGradientUpdater<FloatT, IdxType> up(5, 4, conf, S, b1, b2, eps);
        device_matrix<FloatT> g(4, 2);
        Vec both = kG1; both.insert(both.end(), kG2.begin(), kG2.end());
        to_device(f(both), &g);
        device_matrix<IdxType> idx(1, 6);
        std::vector<long> ids = idsA; ids.insert(ids.end(), idsB.begin(), idsB.end());
        to_device(ids, &idx);
        RepresentationsStorage<FloatT, IdxType>::GradientType desc = {std::forward_as_tuple(g, idx, (size_t)3, nullptr)};
        up.update(st.get(), &desc, lr, lam, S);
        std::vector<Vec> sums; Vec vs;
        adam_state(b1, b2, &sums, &vs);
        Vec m, gmod, table(20, 5.0 * (1 - lam * lr));
        for (int o = 0; o < 5; ++o) for (int k = 0; k < 4; ++k) m.push_back((1 - b1) * sums[o][k]);
        const double bc = std::sqrt(1.0 - b2) / (1.0 - b1);
        for (int x = 0; x < 2; ++x) {
          double av = 0;
          for (int y = 0; y < 3; ++y) av += vs[ids[3 * x + y]];
          for (int k = 0; k < 4; ++k) {
            double am = 0;
            for (int y = 0; y < 3; ++y) am += m[ids[3 * x + y] * 4 + k];
            const double step = bc * (am / 3.0) / (std::sqrt(av / 3.0) + eps);
            gmod.push_back(step);
            for (int y = 0; y < 3; ++y) table[ids[3 * x + y] * 4 + k] += lr * step;
          }
        }
        return near(up.state("m"), m, "m") && near(up.state("v"), vs, "v") && near(to_host(g), gmod, "grad in place") &&
               near(to_host(*st->get()), table, "representations");
      });

      run_case("AdamRepresentationsGradientUpdater DENSE_UPDATE: scalar v per object, dense step" + T, [&] {
        const double eps = 1e-5, b1 = 0.9, b2 = 0.999;
        auto st = constant_storage<RepresentationsStorage<FloatT, IdxType>>(5.0, 5, 4);
        AdamConf conf; conf.set_mode(AdamConf::DENSE_UPDATE);
        AdamRepresentationsGradientUpdater<FloatT, IdxType> up(5, 4, conf, S, b1, b2, eps);
        device_matrix<FloatT> g1(4, 1), g2(4, 1);
        to_device(f(kG1), &g1); to_device(f(kG2), &g2);
        device_matrix<IdxType> i1(1, 3), i2(1, 3);
        to_device(idsA, &i1); to_device(idsB, &i2);
        RepresentationsStorage<FloatT, IdxType>::GradientType desc = {std::forward_as_tuple(g1, i1, (size_t)3, nullptr),
                                                                      std::forward_as_tuple(g2, i2, (size_t)3, nullptr)};
        up.update(st.get(), &desc, lr, lam, S);
        std::vector<Vec> sums; Vec vs;
        adam_state(b1, b2, &sums, &vs);
        const double bc = std::sqrt(1.0 - b2) / (1.0 - b1);
        Vec m, table;
        for (int o = 0; o < 5; ++o)
          for (int k = 0; k < 4; ++k) {
            m.push_back((1 - b1) * sums[o][k]);
            table.push_back(5.0 + lr * (bc * m.back() / (std::sqrt(vs[o]) + eps) - lam * 5.0));
          }
        return near(up.state("m"), m, "m") && near(up.state("v"), vs, "v") && near(to_host(*st->get()), table, "representations");
      });

      run_case("AdamRepresentationsGradientUpdater DENSE_UPDATE_DENSE_VARIANCE (full_adam)" + T, [&] {
        const double eps = 1e-5, b1 = 0.9, b2 = 0.999;
        auto st = constant_storage<RepresentationsStorage<FloatT, IdxType>>(5.0, 5, 4);
        AdamConf conf; conf.set_mode(AdamConf::DENSE_UPDATE_DENSE_VARIANCE);
        AdamRepresentationsGradientUpdater<FloatT, IdxType> up(5, 4, conf, S, b1, b2, eps);
        device_matrix<FloatT> g1(4, 1), g2(4, 1);
        to_device(f(kG1), &g1); to_device(f(kG2), &g2);
        device_matrix<IdxType> i1(1, 3), i2(1, 3);
        to_device(idsA, &i1); to_device(idsB, &i2);
        RepresentationsStorage<FloatT, IdxType>::GradientType desc = {std::forward_as_tuple(g1, i1, (size_t)3, nullptr),
                                                                      std::forward_as_tuple(g2, i2, (size_t)3, nullptr)};
        up.update(st.get(), &desc, lr, lam, S);
        std::vector<Vec> sums; Vec unused;
        adam_state(b1, b2, &sums, &unused);
        const double bc = std::sqrt(1.0 - b2) / (1.0 - b1);
        Vec m, v, table;
        for (int o = 0; o < 5; ++o)
          for (int k = 0; k < 4; ++k) {
            const double g = sums[o][k] - lam * 5.0;   // the L2 term enters the gradient of every element
            m.push_back((1 - b1) * g); v.push_back((1 - b2) * g * g);
            table.push_back(5.0 + lr * (bc * m.back() / (std::sqrt(v.back()) + eps)));
          }
        return near(up.state("m"), m, "m") && near(up.state("v"), v, "v") && near(to_host(*st->get()), table, "representations");
      });
    }

  run_case("Adagrad / sparse Adam refuse several descriptors (the reference aborts)", [&] {
    nvsm_updater* u = nullptr;
    if (nvsm_updater_create(S->ops(), 0, NVSM_ADAGRAD, 0, 10, 4, 0, 0.9f, 0.999f, 1e-6f, &u) != 0) return false;
    device_matrix<FloatT> g(4, 1), table(4, 10);
    device_matrix<IdxType> idx(1, 3);
    nvsm_grad_desc d[2] = {{g.getData(), idx.getData(), 1, 3, nullptr}, {g.getData(), idx.getData(), 1, 3, nullptr}};
    const bool refused = nvsm_updater_update_representations(u, table.getData(), d, 2, 0.1f, 0.0f) != 0;
    const bool negative = nvsm_updater_update_representations(u, table.getData(), d, 1, -0.1f, 0.0f) != 0;   // CHECK_GE(learning_rate, 0)
    nvsm_updater_destroy(u);
    return refused && negative;
  });

  // ---- cpp/model_tests.cu: Representations / Transform -------------------------------------------------------------------
  lse::TrainConfig::UpdateMethodConf sgd;   // type SGD is the default
  auto range_table = [&](RepresentationsStorage<FloatT, IdxType>* r) {   // initialize_range_representations of the reference
    std::vector<FloatT> h(r->num_parameters());
    for (size_t i = 0; i < h.size(); ++i) h[i] = (FloatT)i;
    r->get()->fillwith(nullptr, h);
  };

  run_case("Representations::get_average_representations, plain and weighted (mean divides by the window)", [&] {
    Representations<FloatT, IdxType> reprs(WORD_REPRS, 4, 3, sgd, S);
    RNG rng; reprs.initialize(&rng);
    range_table(&reprs);
    device_matrix<IdxType> idx(6, 1);
    to_device({1L, 3L, 2L, 0L, 3L, 1L}, &idx);
    std::unique_ptr<device_matrix<FloatT>> avg(reprs.get_average_representations(nullptr, idx, 3));
    bool ok = near(to_host(*avg), {(3 + 9 + 6) / 3., (4 + 10 + 7) / 3., (5 + 11 + 8) / 3., (0 + 9 + 3) / 3., (1 + 10 + 4) / 3., (2 + 11 + 5) / 3.}, "plain");
    device_matrix<FloatT> w(6, 1);
    to_device({0.5f, 0.3f, 0.1f, 1.0f, 2.0f, 0.2f}, &w);
    avg.reset(reprs.get_average_representations(nullptr, idx, 3, &w));
    ok = ok && near(to_host(*avg), {(0.5 * 3 + 0.3 * 9 + 0.1 * 6) / 3., (0.5 * 4 + 0.3 * 10 + 0.1 * 7) / 3., (0.5 * 5 + 0.3 * 11 + 0.1 * 8) / 3.,
                                    (1.0 * 0 + 2.0 * 9 + 0.2 * 3) / 3., (1.0 * 1 + 2.0 * 10 + 0.2 * 4) / 3., (1.0 * 2 + 2.0 * 11 + 0.2 * 5) / 3.}, "weighted");
    std::unique_ptr<device_matrix<FloatT>> rows(reprs.get_representations(nullptr, idx));
    std::unique_ptr<device_matrix<FloatT>> one(reprs.get_representation(2));
    return ok && near(to_host(*rows), {3, 4, 5, 9, 10, 11, 6, 7, 8, 0, 1, 2, 9, 10, 11, 3, 4, 5}, "get_representations") &&
           near(to_host(*one), {6, 7, 8}, "get_representation") && reprs.initialized() && reprs.num_objects() == 4 && reprs.size() == 3;
  });

  run_case("Representations::update: dense decay with a zero gradient, scatter with window 2", [&] {
    Representations<FloatT, IdxType> reprs(WORD_REPRS, 4, 3, sgd, S);
    range_table(&reprs);
    device_matrix<IdxType> words(1, 4);
    to_device({0L, 3L, 1L, 0L}, &words);
    device_matrix<FloatT> ww(1, 4);
    ww.fillwith(nullptr, 1.0f);
    device_matrix<FloatT> g(3, 2);   // zero
    Representations<FloatT, IdxType>::GradientType desc = {std::forward_as_tuple(g, words, (size_t)2, &ww)};
    reprs.update(&desc, 0.1f, 0.1f / 2.0f, S);   // scaled lambda = lambda / batch (cpp/intermediate_results.cu:126-129)
    const double s = 1.0 - (0.1 * 0.1) / 2.0;
    Vec want(12);
    for (int i = 0; i < 12; ++i) want[i] = i * s;
    bool ok = near(to_host(*reprs.get()), want, "decay only");
    range_table(&reprs);
    to_device({5.0f, 4.0f, 3.0f, -3.0f, -2.0f, 10.0f}, &g);
    reprs.update(&desc, 0.1f, 0.0f, S);
    ok = ok && near(to_host(*reprs.get()), {0. + (5.0 - 3.0) * 0.1, 1. + (4.0 - 2.0) * 0.1, 2. + (3.0 + 10.0) * 0.1, 3. - 3.0 * 0.1, 4. - 2.0 * 0.1,
                                            5. + 10.0 * 0.1, 6., 7., 8., 9. + 5.0 * 0.1, 10. + 4.0 * 0.1, 11. + 3.0 * 0.1}, "scatter");
    reprs.update(nullptr, 0.1f, 0.0f, S);   // "No gradient": nothing happens
    return ok && std::fabs(reprs.get_parameter_gradient(desc, 0) - (5.0 - 3.0)) < 1e-6 && std::fabs(reprs.get_parameter_gradient(desc, 11) - 3.0) < 1e-6;
  });

  run_case("RepresentationsStorage::update_dense", [&] {
    RepresentationsStorage<FloatT, IdxType> reprs(4, 3, S);
    range_table(&reprs);
    device_matrix<FloatT> g(3, 4);
    g.fillwith(nullptr, 10.0f);
    reprs.update_dense(nullptr, g, 0.1f, 0.01f);
    Vec want(12);
    for (int i = 0; i < 12; ++i) want[i] = i * (1.0 - 0.01 * 0.1) + 10.0 * 0.1;
    return near(to_host(*reprs.get()), want, "dense") && reprs.get_data().count("representations") == 1 && reprs.num_parameters() == 12;
  });

  auto counting_transform = [&](Transform<FloatT>* t) {   // initialize_transform of the reference: 0..14, bias 0..4 * 1e-3
    std::vector<FloatT> h(15), b(5);
    for (int i = 0; i < 15; ++i) h[i] = (FloatT)i;
    for (int i = 0; i < 5; ++i) b[i] = (FloatT)(i * 1e-3);
    std::get<0>(t->get())->fillwith(nullptr, h);
    std::get<1>(t->get())->fillwith(nullptr, b);
  };

  run_case("Transform::transform: tanh(T p + b), closed form", [&] {
    lse::ModelDesc::TransformDesc desc;
    Transform<FloatT> t(TRANSFORM, desc, 3, 5, sgd, S);
    RNG rng; t.initialize(&rng);
    counting_transform(&t);
    device_matrix<FloatT> in(3, 2);
    to_device({0.01f, 0.02f, 0.03f, 0.001f, 0.002f, 0.003f}, &in);
    std::unique_ptr<device_matrix<FloatT>> out(t.transform(nullptr, in, nullptr));
    Vec want;
    for (double z : {0.400, 0.461, 0.522, 0.583, 0.644, 0.040, 0.047, 0.054, 0.061, 0.068}) want.push_back(std::tanh(z));
    return near(to_host(*out), want, "transform") && t.source_repr_size() == 3 && t.target_repr_size() == 5 && t.num_parameters() == 20;
  });

  run_case("Transform::transform through BatchNormalization + tanh: the reference's golden values", [&] {
    BatchNormalization<FloatT> bn(5, 0.1f, 1e-5f, true);
    lse::ModelDesc::TransformDesc desc;
    Transform<FloatT> t(TRANSFORM, desc, 3, 5, sgd, S);
    counting_transform(&t);
    device_matrix<FloatT> in(3, 2);
    to_device({0.01f, 0.02f, 0.03f, 0.001f, 0.002f, 0.003f}, &in);
    std::unique_ptr<device_matrix<FloatT>> out(t.transform(nullptr, in, &bn));
    return near(to_host(*out), {0.7615293524851600715, 0.76196488305628828908, 0.76239459573842593976, 0.76282051982955900726,
                                0.76324378068549525445, -0.76152935248515996047, -0.76112478489184165475, -0.76071446308133705561,
                                -0.76030038648736808504, -0.75988366421371911219}, "bn + tanh", 2e-5);
  });

  run_case("Glorot initialisation consumes the shared engine like Model::initialize (W, then T; b = 0)", [&] {
    // Representations::initialize + Transform::initialize from one RNG == the first and third tensor of nvsm_initialize
    lse::ModelDesc md; md.set_word_repr_size(6); md.set_entity_repr_size(4);
    lse::TrainConfig tc; tc.set_batch_size(8); tc.set_window_size(2); tc.set_num_random_entities(1);
    DefaultModel model(7, 5, md, tc, 0, NVSM_GEMM_FP32);
    RNG a; a.seed(11); RNG b; b.seed(11);
    model.initialize(&a);
    Representations<FloatT, IdxType> W(WORD_REPRS, 7, 6, sgd, S), E(ENTITY_REPRS, 5, 4, sgd, S);
    Transform<FloatT> T(TRANSFORM, md.transform_desc(), 6, 4, sgd, S);
    W.initialize(&b); E.initialize(&b); T.initialize(&b);
    const auto data = model.get_data();
    const Vec wantW(data.at("word_representations-representations").data.begin(), data.at("word_representations-representations").data.end());
    const Vec wantE(data.at("entity_representations-representations").data.begin(), data.at("entity_representations-representations").data.end());
    const Vec wantT(data.at("word_entity_mapping-transform").data.begin(), data.at("word_entity_mapping-transform").data.end());
    return near(to_host(*W.get()), wantW, "W", 0, 0) && near(to_host(*E.get()), wantE, "E", 0, 0) &&
           near(to_host(*std::get<0>(T.get())), wantT, "T", 0, 0) && nvsm_detail::rng_get_state(a) == nvsm_detail::rng_get_state(b);
  });

  // ---- cpp/cudnn_utils_tests.cu ---------------------------------------------------------------------------------------------
  run_case("BatchNormalization: constant input normalises to zero", [&] {
    BatchNormalization<FloatT> bn(10);
    device_matrix<FloatT> in(10, 100), bias(10, 1);
    in.fillwith(nullptr, 1.0f);
    bn.forward(in, bias, &in);
    return near(to_host(in), Vec(1000, 0.0), "zeros", 0, 1e-6);
  });

  run_case("BatchNormalization forward (closed form) and backward (golden grad_bias, dx vs float64)", [&] {
    const double eps = 1e-5;
    BatchNormalization<FloatT> bn(3, 0.1f, (FloatT)eps);
    device_matrix<FloatT> bias(3, 1), in(3, 2), out(3, 2);
    in.fillwith(nullptr, std::vector<FloatT>{1.0f, 2.0f, 3.0f, 5.0f, 10.0f, 20.0f});
    bn.forward(in, bias, &out);
    bool ok = near(to_host(out), {(1.0 - 3.0) / std::sqrt(4.0 + eps), (2.0 - 6.0) / std::sqrt(16.0 + eps), (3.0 - 11.5) / std::sqrt(72.25 + eps),
                                  (5.0 - 3.0) / std::sqrt(4.0 + eps), (10.0 - 6.0) / std::sqrt(16.0 + eps), (20.0 - 11.5) / std::sqrt(72.25 + eps)}, "forward");
    device_matrix<FloatT> grad(3, 2), gbias(3, 1);
    grad.fillwith(nullptr, std::vector<FloatT>{0.25f, -0.1f, 0.3f, 1.0f, 0.005f, -0.5f});
    bn.backward(grad, in, bias, &grad, &gbias);
    ok = ok && near(to_host(gbias), {1.25, -0.095, -0.2}, "grad_bias");
    // dx is the difference of nearly equal terms here (two instances): float32 keeps ~1e-7 absolute of the reference's
    // float64 literals -4.6874824e-07, -8.2031173e-09, 6.5133306e-09, ...
    return ok && near(to_host(grad), {-4.687482422216504574e-07, -8.2031173104235577398e-09, 6.5133306248466027455e-09, 4.687482422216504574e-07,
                                      8.2031173086888342638e-09, -6.5133306248466027455e-09}, "grad_input", 0, 1.5e-7);
  });

  run_case("BatchNormalization backward on a well-conditioned batch vs a float64 restatement; in place == out of place", [&] {
    const int C = 256, N = 4096;
    std::vector<FloatT> x((size_t)C * N), dy((size_t)C * N);
    unsigned s = 12345;
    auto rnd = [&] { s = s * 1664525u + 1013904223u; return (FloatT)((s >> 8) & 0xFFFF) / 65536.0f - 0.5f; };
    for (auto& v : x) v = 3.0f * rnd() + 0.25f;
    for (auto& v : dy) v = rnd();
    BatchNormalization<FloatT> bn1(C), bn2(C);
    device_matrix<FloatT> in(C, N), in2(C, N), out(C, N), bias(C, 1), g(C, N), g2(C, N), dx(C, N), gb1(C, 1), gb2(C, 1);
    in.fillwith(nullptr, x); in2.fillwith(nullptr, x); g.fillwith(nullptr, dy); g2.fillwith(nullptr, dy);
    std::unique_ptr<device_matrix<FloatT>> keep(in.copy());
    bn1.forward(in, bias, &in);            // in place
    bn2.forward(in2, bias, &out);          // out of place
    if (to_host(in) != to_host(out)) return false;
    bn1.backward(g, *keep, bias, &g, &gb1);
    bn2.backward(g2, *keep, bias, &dx, &gb2);
    if (to_host(g) != to_host(dx) || to_host(gb1) != to_host(gb2)) return false;
    Vec want((size_t)C * N), wantb(C);
    for (int c = 0; c < C; ++c) {
      double mu = 0, var = 0, sb = 0, sg = 0;
      for (int i = 0; i < N; ++i) mu += x[(size_t)i * C + c];
      mu /= N;
      for (int i = 0; i < N; ++i) var += (x[(size_t)i * C + c] - mu) * (x[(size_t)i * C + c] - mu);
      const double is = 1.0 / std::sqrt(var / N + 1e-4);
      for (int i = 0; i < N; ++i) { const double xh = (x[(size_t)i * C + c] - mu) * is; sb += dy[(size_t)i * C + c]; sg += dy[(size_t)i * C + c] * xh; }
      wantb[c] = sb;
      for (int i = 0; i < N; ++i) {
        const double xh = (x[(size_t)i * C + c] - mu) * is;
        want[(size_t)i * C + c] = is * (dy[(size_t)i * C + c] - sb / N - xh * sg / N);
      }
    }
    return near(to_host(dx), want, "dx", 1e-4, 1e-5) && near(to_host(gb2), wantb, "grad_bias", 1e-4, 1e-4);
  });

  S->synchronize();
  std::printf("%d cases, %d failed, %ld kernel launches\n", g_cases, g_failed, nvsm_ops_kernel_launches(S->ops()));
  if (g_failed == 0) std::printf("CLASSES_TEST_OK\n");
  return g_failed;
}
