// Overlay for the reference build "cfg2" (BASELINE.json configs[1]): one blood-cell type scaled to
// 100 000 particles.  Written in the reference's own config format (src/config/blood_cells_definition.hpp);
// presets come from the reference's blood_cell_presets.hpp in the scratch copy.
#pragma once

#include "blood_cell_presets.hpp"
#include "../meta_factory/blood_cells_def_type.hpp"

#include <boost/mp11/list.hpp>

namespace
{
	using namespace boost::mp11;

	using UserDefinedBloodCellList = mp_list<
	BloodCellDef<5000, 20, 108, 15720158,
		preset::White_blood_cell_One_Springs,
		preset::White_blood_cell_One_Vertices,
		preset::White_blood_cell_One_Indices,
		preset::White_blood_cell_One_Normals>
	>;
}
