// The reference's OWN distributed-tensor test and timing program (tests/dist.cpp of
// eromero-vlc/superbblas, included from where it lies under /root/reference -- never copied),
// compiled UNCHANGED against the drop-in header include/superbblas.h.  It checks the partition
// generators and make_hole against its known answers, then times permuting copies, periodic shifts,
// halo fills, batched GEMMs (detail::xgemm_batch_strided) and contractions of "xyztscn" tensors with
// host (CPU-context, staged through the GPU) and with GPU components.
//
//   ref_dist_wrapper [--dim='x y z t n'] [--reps=r]
#define SUPERBBLAS_USE_GPU
#include "tests/dist.cpp"
