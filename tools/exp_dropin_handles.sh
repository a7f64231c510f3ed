# distribution of the multi-handle staged modes (5 repetitions each)
for c in device,ref,2,-1,4 device,ref,4,-1,4 device,ref,2,-1,2 device,ref,2,-1,8 device,ref,2,-1,1 device,ref,0,-1,4 device,pageable,2,-1,4 device,pageable,2,-1,8 device,pageable,2,-1,1 device,pageable,0,-1,4 device,pinned,2,-1,4; do
  for rep in 1 2 3 4 5; do echo "$(tools/jm_dropin --frames 400 --custom $c)"; done
done
