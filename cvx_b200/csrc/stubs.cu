// entry points of include/cvxb.h whose device path is not built yet
#include "kkt.cuh"
#define NOTIMPL(name) cvxb::set_last_error(name ": not built yet"); return CVXB_ENOTIMPL
extern "C" {
int cvxb_batch_create(cvxb_handle, const cvxb_batch_desc*, cvxb_batch*) { NOTIMPL("cvxb_batch_create"); }
int cvxb_batch_destroy(cvxb_batch) { return CVXB_OK; }
int cvxb_batch_barrier_solve(cvxb_handle, cvxb_batch, const cvxb_params*, cvxb_batch_result*) { NOTIMPL("cvxb_batch_barrier_solve"); }
}
