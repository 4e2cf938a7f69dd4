#include "map2d_b200.h"
#include <stdio.h>
int main(void) {
    m2d_config c;
    m2d_handle h = NULL;
    double tl[3], br[3], plane[7] = {0, 0, 0, 0, 0, 0, 1}, org[2] = {108.0, 34.0};
    m2d_config_default(&c);
    int rc = m2d_create(M2D_TYPE_MULTIBAND, &c, &h);
    printf("create rc=%d handle=%p classes=%d\n", rc, (void*)h, M2D_KERNEL_CLASSES);
    if (m2d_tile_gps_corners(plane, 0, 0, 25.6, 0, 0, org, tl, br) != M2D_OK) return 2;
    printf("%.9f %.9f\n", br[0], br[1]);
    if (h) m2d_destroy(h);
    return 0;
}
