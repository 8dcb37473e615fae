#include <stdio.h>
#include "../include/msckf_b200_presets.h"
int main(void) {
    printf("%zu %zu %zu %zu %zu %zu\n", sizeof(mskf_config), sizeof(mskf_feature), sizeof(mskf_tracking_info),
           sizeof(mskf_grid_feature), sizeof(mskf_state), sizeof(mskf_cam_state));
    return 0;
}
