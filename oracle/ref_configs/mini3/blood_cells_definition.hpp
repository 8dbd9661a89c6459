// Overlay for the reference build "mini3": a small mixed scene that exercises the corners of the
// meta-factory and of the per-type launch logic:
//   * a duplicated user definition (White_blood_cell twice)  -> fold + mp_unique_if   (blood_cell_factory.hpp:60-115)
//   * a power-of-two type (8 particles per cell, authored here) -> comparator ordering (:119-162) and the
//     warp-sync vein-end variant (vein_end.cu:23-30,111-138)
//   * three types with different model offsets                 -> the per-slice radius lookup (particle_collisions.cuh:76,124)
// Reference config format; the two 20-particle presets come from the reference's blood_cell_presets.hpp.
#pragma once

#include "blood_cell_presets.hpp"
#include "../meta_factory/blood_cells_def_type.hpp"

#include <boost/mp11/list.hpp>

namespace preset
{
	using namespace boost::mp11;

	// An 8-vertex box cell, edge 3, all 12 edges + 4 space diagonals as springs.
	using Box8_Vertices = mp_list<
		mp_float3<-1500000, -1500000, -1500000>,
		mp_float3< 1500000, -1500000, -1500000>,
		mp_float3< 1500000,  1500000, -1500000>,
		mp_float3<-1500000,  1500000, -1500000>,
		mp_float3<-1500000, -1500000,  1500000>,
		mp_float3< 1500000, -1500000,  1500000>,
		mp_float3< 1500000,  1500000,  1500000>,
		mp_float3<-1500000,  1500000,  1500000>
	>;
	using Box8_Springs = mp_list<
		Spring<0, 1, 3000000>, Spring<1, 2, 3000000>, Spring<2, 3, 3000000>, Spring<3, 0, 3000000>,
		Spring<4, 5, 3000000>, Spring<5, 6, 3000000>, Spring<6, 7, 3000000>, Spring<7, 4, 3000000>,
		Spring<0, 4, 3000000>, Spring<1, 5, 3000000>, Spring<2, 6, 3000000>, Spring<3, 7, 3000000>,
		Spring<0, 6, 5196152>, Spring<1, 7, 5196152>, Spring<2, 4, 5196152>, Spring<3, 5, 5196152>
	>;
	using Box8_Indices = mp_list<mp_int<0>, mp_int<1>, mp_int<2>>;
	using Box8_Normals = mp_list<mp_float3<0, 0, 1000000>>;
}

namespace
{
	using namespace boost::mp11;

	using UserDefinedBloodCellList = mp_list<
	BloodCellDef<30, 20, 108, 15720158,
		preset::White_blood_cell_One_Springs,
		preset::White_blood_cell_One_Vertices,
		preset::White_blood_cell_One_Indices,
		preset::White_blood_cell_One_Normals>,

	BloodCellDef<40, 8, 3, 255,
		preset::Box8_Springs,
		preset::Box8_Vertices,
		preset::Box8_Indices,
		preset::Box8_Normals>,

	BloodCellDef<50, 20, 108, 14352898,
		preset::Blood_dust_One_Springs,
		preset::Blood_dust_One_Vertices,
		preset::Blood_dust_One_Indices,
		preset::Blood_dust_One_Normals>,

	BloodCellDef<20, 20, 108, 15720158,
		preset::White_blood_cell_One_Springs,
		preset::White_blood_cell_One_Vertices,
		preset::White_blood_cell_One_Indices,
		preset::White_blood_cell_One_Normals>
	>;
}
