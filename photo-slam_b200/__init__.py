"""photo-slam_b200 — B200-native (sm_100a) drop-in for Photo-SLAM's Gaussian-splatting hot path.

Host-side mirror of the reference's operator interface for this path:
  rasterizer.GaussianRasterizationSettings / GaussianRasterizer / rasterize_gaussians
      (reference include/gaussian_rasterizer.h:25-127, src/gaussian_rasterizer.cpp)
  rasterizer.RasterizeGaussiansCUDA / RasterizeGaussiansBackwardCUDA / markVisible
      (reference include/rasterize_points.h:18-65)
All compute goes through the C-ABI of lib/libpsb200.so (include/psb200.h); there is no CPU fallback:
importing the ops without the built library raises.
"""
__version__ = "0.1.0"
