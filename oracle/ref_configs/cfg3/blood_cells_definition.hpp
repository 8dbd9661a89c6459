// Overlay for the reference build "cfg3" (BASELINE.json configs[2]): mixed blood-cell types at
// 1 000 000 particles.  Reference config format; presets from the reference's blood_cell_presets.hpp.
#pragma once

#include "blood_cell_presets.hpp"
#include "../meta_factory/blood_cells_def_type.hpp"

#include <boost/mp11/list.hpp>

namespace
{
	using namespace boost::mp11;

	using UserDefinedBloodCellList = mp_list<
	BloodCellDef<25000, 20, 108, 15720158,
		preset::White_blood_cell_One_Springs,
		preset::White_blood_cell_One_Vertices,
		preset::White_blood_cell_One_Indices,
		preset::White_blood_cell_One_Normals>,

	BloodCellDef<25000, 20, 108, 14352898,
		preset::Blood_dust_One_Springs,
		preset::Blood_dust_One_Vertices,
		preset::Blood_dust_One_Indices,
		preset::Blood_dust_One_Normals>
	>;
}
