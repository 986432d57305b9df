// utils/vlp_common.h, utils/pcl_utils.h — licalib::VPoint / TPoint and their clouds (L/include/utils/pcl_utils.h:39-70)
#include "../pcl/pcl_b200.h"
