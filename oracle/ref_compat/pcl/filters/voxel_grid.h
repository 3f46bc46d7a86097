// oracle/ref_compat/pcl/filters/voxel_grid.h -- TEST INFRASTRUCTURE ONLY: src/projection.cpp includes the PCL header without
// using it.
#pragma once
