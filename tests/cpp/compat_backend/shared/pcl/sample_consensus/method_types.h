// TEST INFRASTRUCTURE ONLY: forwards to the PCL stand-in (oracle/ref_compat/pcl_standin.h).
#pragma once
#include "../../../../../../oracle/ref_compat/pcl_standin.h"
