// Stand-in for <gtest/gtest_prod.h>: the reference grants its own tests access to internals through
// FRIEND_TEST; oracle/ref_shim/ref_capi.cu reads gradients / fills batches through the same door.
#ifndef REF_SHIM_GTEST_PROD_H
#define REF_SHIM_GTEST_PROD_H
#define FRIEND_TEST(test_case_name, test_name) friend class test_case_name##_##test_name##_Test
#endif
