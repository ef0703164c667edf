// TEST INFRASTRUCTURE (fixture generator, compiled by tests/golden/make_hctest_fixture.sh against the objects of the UNMODIFIED
// reference build; never part of the product).
//
// Replays an hctest snapshot through the real reference -- Nyx::integrate_state_struct with nyx.hctest_example_read = 1, the call
// Exec/HeatCoolTests makes through Nyx::advance_heatcool (Source/TimeStep/Nyx_advance.cpp:438-527) -- and writes what that driver
// throws away: the post-step S_new, D_old (= diag_eos) and IR of every box, as FABs in `<chunk prefix>out.<box>`.
// The MultiFab set-up below follows advance_heatcool:487-516 line by line (same ghost widths, same setVal calls, sdc_iter = 0);
// the shipped driver itself cannot be used: it casts the Nyx level to an Amr (Exec/HeatCoolTests/nyx_main.cpp:104).
#include <AMReX.H>
#include <AMReX_MultiFab.H>
#include <AMReX_ParmParse.H>
#include <Nyx.H>

#include <fstream>
#include <string>

using namespace amrex;

std::string inputs_name = "";            // the reference's Nyx_output.cpp refers to it (Exec/*/nyx_main.cpp define it)
amrex::LevelBld* getLevelBld();

int main(int argc, char* argv[])
{
    amrex::Initialize(argc, argv);
    {
        if (argc > 1) inputs_name = argv[1];
        Nyx* level = new Nyx();
        level->variable_setup();

        ParmParse pp_nyx("nyx");
        int index = 0;
        pp_nyx.query("hctest_example_index", index);
        std::string f_badmap = "hctest/BADMAP." + std::to_string(index);
        std::string f_chunk = "hctest/Chunk." + std::to_string(index) + ".";
        pp_nyx.query("hctest_filename_badmap", f_badmap);
        pp_nyx.query("hctest_filename_chunk", f_chunk);

        BoxArray grids;
        DistributionMapping dmap;
        {
            std::ifstream ifs(f_badmap.c_str());
            grids.readFrom(ifs);
            dmap.readFrom(ifs);
        }
        const int NUM_STATE = Nyx::NUM_STATE;   // NUM_GROW is the macro of Source/Hydro/IndexDefines.H:39
        MultiFab S_new(grids, dmap, NUM_STATE, NUM_GROW);
        MultiFab D_old(grids, dmap, 2, NUM_GROW);
        MultiFab IR_old(grids, dmap, 1, 0);
        MultiFab S_old_tmp(grids, dmap, NUM_STATE, NUM_GROW);
        S_old_tmp.setVal(0.);
        MultiFab hydro_src(grids, dmap, NUM_STATE, 0);
        hydro_src.setVal(0.);
        MultiFab reset_e_src(grids, dmap, 1, NUM_GROW);
        reset_e_src.setVal(0.0);

        Real initial_z = 0, final_z = 0, fixed_dt = 0;
        pp_nyx.get("initial_z", initial_z);
        pp_nyx.get("final_z", final_z);
        pp_nyx.get("fixed_dt", fixed_dt);
        const Real a = 1 / (initial_z + 1), a_end = 1 / (final_z + 1);
        level->integrate_state_struct(S_old_tmp, S_new, D_old, hydro_src, IR_old, reset_e_src, a, a_end, fixed_dt, 0);

        for (MFIter mfi(S_new); mfi.isValid(); ++mfi) {
            std::ofstream ofs((f_chunk + "out." + std::to_string(mfi.index())).c_str());
            S_new[mfi].writeOn(ofs);
            D_old[mfi].writeOn(ofs);
            IR_old[mfi].writeOn(ofs);
        }
        amrex::Print() << "replayed " << grids.size() << " boxes: a = " << a << " a_end = " << a_end << " dt = " << fixed_dt << "\n";
    }
    amrex::Finalize();
    return 0;
}
