// backend.cu — placeholder while the front end is brought up (replaced by the fp64 filter).
#include "common.cuh"
struct BeBuffers { int dummy; };
int be_create(mskf_handle *h) { h->bb = new BeBuffers; return MSKF_OK; }
void be_destroy(mskf_handle *h) { delete h->bb; h->bb = nullptr; }
int be_step(mskf_handle *h, const std::vector<int> &, const mskf_feature *, int, int, double) { h->err = "back end not built"; return MSKF_ERR_STATE; }
int be_init_gravity(mskf_handle *, int) { return MSKF_OK; }
int be_get_state(mskf_handle *h, int, mskf_state *) { h->err = "back end not built"; return MSKF_ERR_STATE; }
int be_get_cam_states(mskf_handle *h, int, mskf_cam_state *, int, int *) { h->err = "back end not built"; return MSKF_ERR_STATE; }
int be_get_cov(mskf_handle *h, int, double *, int, int *) { h->err = "back end not built"; return MSKF_ERR_STATE; }
int be_reset(mskf_handle *, int) { return MSKF_OK; }
