for rep in 1 2 3; do
for v in device_surface_ref_delay2_4_handles_fps device_surface_pageable_out_delay2_4_handles_fps device_surface_pinned_out_delay2_4_handles_fps device_surface_pageable_out_delay2_fps device_surface_ref_delay2_fps; do
  echo "default $(tools/jm_dropin --frames 400 --only $v)"
  echo "conn32  $(CUDA_DEVICE_MAX_CONNECTIONS=32 tools/jm_dropin --frames 400 --only $v)"
done; done
