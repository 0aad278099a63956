/*
 * ref_chem.cpp -- parity harness around the reference's finite-rate chemistry
 * (ucs/chem.tcc, reaction.tcc, species.tcc).  TEST INFRASTRUCTURE ONLY; no chemistry
 * arithmetic of its own.
 *
 *   ref_chem <rxn case (path without .rxn)> <species table> <states.bin> <outdir>
 *
 * 1. writes <outdir>/chemdb.hdf5 in the schema Species::GetDBInfo reads (species.tcc:140-322) from the
 *    species table (tools/make_chem_golden.py extracts it from the reference's own chemdata/BURCAT_FIXED.THR),
 *    using the reference's HDF layer;
 * 2. constructs the reference's ChemModel<Real> from the reference's .rxn file + that database;
 * 3. dumps the model tables as the reference parsed them (species MW / NASA-7 ranges, reaction constants,
 *    stoichiometry, third-body efficiencies) and, for every state (rho_i [kg/m^3], T [K]) in states.bin,
 *    ChemModel::GetMassProductionRates (chem.tcc:575-583).
 */
#include <cmath>
#include <complex>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <iostream>
#include <limits>
#include <sstream>
#include <string>
#include <vector>
#include <sys/stat.h>
#include <mpi.h>

#define private public
#define protected public
#include "general.h"
#include "exceptions.h"
#include "h5layer.h"
#include "chem.h"
#include "reaction.h"
#include "species.h"
#undef private
#undef protected

template <class T>
static void Dump(const std::string& dir, const std::string& name, const T* data, size_t n)
{
  std::string fn = dir + "/" + name + ".bin";
  FILE* f = fopen(fn.c_str(), "wb");
  if(!f){ perror(fn.c_str()); exit(2); }
  if(n) fwrite(data, sizeof(T), n, f);
  fclose(f);
}

int main(int argc, char* argv[])
{
  MPI_Init(&argc, &argv);
  if(argc < 5){
    std::cerr << "usage: " << argv[0] << " <rxn case> <species table> <states.bin> <outdir>" << std::endl;
    return 1;
  }
  std::string rxncase = argv[1], table = argv[2], states = argv[3], out = argv[4];
  mkdir(out.c_str(), 0777);
  Abort.rootDirectory = out + "/";
  freopen((out + "/log.txt").c_str(), "w", stdout);

  HDF_TurnOffErrorHandling();
  // ---- 1. species database
  std::string db = out + "/chemdb.hdf5";
  {
    std::ifstream fin(table.c_str());
    hid_t h5 = HDF_OpenFile(db, 1);
    if(h5 < 0){ std::cerr << "cannot create " << db << std::endl; return 2; }
    std::string sym;
    while(fin >> sym){
      Real mw, hf, lo[7], hi[7];
      fin >> mw >> hf;
      for(int i = 0; i < 7; i++) fin >> lo[i];
      for(int i = 0; i < 7; i++) fin >> hi[i];
      std::string dir = "/species/" + sym;
      HDF_WriteScalar(h5, dir, "MW", &mw);
      HDF_WriteScalar(h5, dir, "NASA7_burcat_coeff15", &hf);
      HDF_WriteArray(h5, dir, "NASA7_burcat1", lo, 7);
      HDF_WriteArray(h5, dir, "NASA7_burcat2", hi, 7);
      // NASA RP-1311 transport fits (rows of [Tlo, Thi, A, B, C, D]) from the reference's chemdata/trans.inp
      int nk = 0, nmu = 0;
      Real kfit[18], mufit[18];
      fin >> nk;
      if(nk < 1 || nk > 3){ std::cerr << "bad conductivity table for " << sym << std::endl; return 2; }
      for(int i = 0; i < nk*6; i++) fin >> kfit[i];
      fin >> nmu;
      if(nmu < 1 || nmu > 3){ std::cerr << "bad viscosity table for " << sym << std::endl; return 2; }
      for(int i = 0; i < nmu*6; i++) fin >> mufit[i];
      HDF_WriteArray(h5, dir, "k", kfit, nk, 6);
      HDF_WriteArray(h5, dir, "mu", mufit, nmu, 6);
    }
    HDF_CloseFile(h5);
  }

  // ---- 2. the reference's chemistry model
  ChemModel<Real> chem(rxncase, db);
  Int ns = chem.nspecies, nr = chem.nreactions;

  // ---- 3. tables
  {
    std::vector<Real> mw(ns), coeff(ns*14);
    for(Int i = 0; i < ns; i++){
      mw[i] = chem.species[i].MW;
      for(Int k = 0; k < 7; k++){
	coeff[i*14 + k] = chem.species[i].thermo_coeff[0][k];
	coeff[i*14 + 7 + k] = chem.species[i].thermo_coeff[1][k];
      }
    }
    {
      std::ofstream fn((out + "/species_names.txt").c_str());
      for(Int i = 0; i < ns; i++) fn << chem.species[i].symbol << "\n";
    }
    Dump(out, "species_mw", mw.data(), mw.size());
    Dump(out, "species_nasa7", coeff.data(), coeff.size());
    std::vector<Real> rk(nr*3), nup(nr*ns, 0.0), nupp(nr*ns, 0.0), tbeff(nr*ns, 1.0);
    std::vector<Int> flags(nr*4), inrxn(nr*ns, 0), order(nr*ns, -1);
    for(Int j = 0; j < nr; j++){
      Reaction<Real>& r = chem.reactions[j];
      rk[j*3] = r.A; rk[j*3+1] = r.EA; rk[j*3+2] = r.n;
      flags[j*4] = r.rxnType; flags[j*4+1] = r.thirdBodiesPresent; flags[j*4+2] = r.backwardRateGiven;
      flags[j*4+3] = r.GetNspecies();
      for(Int k = 0; k < r.GetNspecies(); k++){
	Int g = r.globalIndx[k];
	order[j*ns + k] = g;                 // local -> global, in the reaction's own species order
	inrxn[j*ns + g] = 1;
	nup[j*ns + k] = r.Nup[k];
	nupp[j*ns + k] = r.Nupp[k];
	if(r.thirdBodiesPresent && (size_t)k < r.TBEff.size()) tbeff[j*ns + k] = r.TBEff[k];
      }
    }
    Dump(out, "rxn_A_EA_n", rk.data(), rk.size());
    Dump(out, "rxn_flags", flags.data(), flags.size());
    Dump(out, "rxn_species", order.data(), order.size());
    Dump(out, "rxn_nup", nup.data(), nup.size());
    Dump(out, "rxn_nupp", nupp.data(), nupp.size());
    Dump(out, "rxn_tbeff", tbeff.data(), tbeff.size());
    Int dims[2] = {ns, nr};
    Dump(out, "dims", dims, 2);
  }

  // ---- 4. mass production rates on the given states: records of ns+1 doubles (rho_i ..., T)
  {
    FILE* f = fopen(states.c_str(), "rb");
    if(!f){ perror(states.c_str()); return 2; }
    fseek(f, 0, SEEK_END);
    size_t n = ftell(f)/sizeof(Real)/(ns + 1);
    fseek(f, 0, SEEK_SET);
    std::vector<Real> st(n*(ns + 1)), wdot(n*ns), kf(n*nr), kb(n*nr);
    if(fread(st.data(), sizeof(Real), st.size(), f) != st.size()){ std::cerr << "short read" << std::endl; return 2; }
    fclose(f);
    for(size_t s = 0; s < n; s++){
      Real* rhoi = &st[s*(ns + 1)];
      Real T = rhoi[ns];
      chem.GetMassProductionRates(rhoi, T, &wdot[s*ns]);
      for(Int j = 0; j < nr; j++){
	kf[s*nr + j] = chem.reactions[j].GetForwardReactionRate(T);
	kb[s*nr + j] = chem.reactions[j].GetBackwardReactionRate(T);
      }
    }
    Dump(out, "wdot", wdot.data(), wdot.size());
    Dump(out, "kf", kf.data(), kf.size());
    Dump(out, "kb", kb.data(), kb.size());
  }
  fflush(stdout);
  MPI_Finalize();
  H5close();   // the model keeps the database open; flush it so that <outdir>/chemdb.hdf5 is a valid file for ref_harness
  _exit(0);
}
