/* Shadows the reference's io/map_saver.hpp (TEST INFRASTRUCTURE ONLY).
 * loop_detector_{branch_bound,grid_search}.cpp include it but use nothing
 * from it; the real header drags in boost::gil / libpng, absent here. */
#pragma once
