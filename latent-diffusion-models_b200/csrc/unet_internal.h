// Private cross-TU interface of the UNet handle (not part of the C ABI).
#pragma once
#include <stdint.h>
struct ldm_unet;
// ldm_unet_forward with the input rows taken modulo x_batch: rows [0,batch) read image (row % x_batch).
// Lets the sampler run cond+uncond as one 2B-row pass over a B-row x_t without duplicating it.
int ldm_unet_forward_ex(ldm_unet* h, const float* x, int x_batch, const int64_t* t, const int64_t* t_dev_scalar,
                        const int64_t* y, int y_len, int y_rows, int batch, float* out, void* workspace,
                        int64_t workspace_bytes, void* stream);
int ldm_unet_in_channels(const ldm_unet* h);
int ldm_unet_out_channels(const ldm_unet* h);
int ldm_unet_image_size(const ldm_unet* h);
