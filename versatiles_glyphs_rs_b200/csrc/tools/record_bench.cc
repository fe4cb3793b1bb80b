// record_bench.cc — host-side cost of recording outlines (GlyphBatch::add_glyph with the dummy renderer: cmap, glyf walk,
// curve records, exact frame) over the 20 Noto Sans fixture files, in microseconds per glyph.
// g++ -O2 -std=c++17 -I include tools/record_bench.cc -L <libdir> -lvgb200host -lb200sdf -Wl,-rpath,<libdir>
#include "vgb200_host.h"
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <string>
#include <dirent.h>
#include <algorithm>
int main(int argc, char**argv){
  std::string dir = std::string(getenv("VGB_REPO") ? getenv("VGB_REPO") : "/root/repo") + "/testdata/Noto Sans/";
  std::vector<std::string> names; DIR*d=opendir(dir.c_str()); while(auto e=readdir(d)){std::string n=e->d_name; if(n.size()>4&&n.substr(n.size()-4)==".ttf")names.push_back(n);} closedir(d); std::sort(names.begin(),names.end());
  std::vector<vgb_font*> fonts; std::vector<std::vector<uint32_t>> cps;
  for(auto&n:names){ auto f=vgb_font_from_path((dir+n).c_str()); fonts.push_back(f); std::vector<uint32_t> c(70000); size_t k=vgb_font_codepoints(f,c.data(),c.size()); c.resize(k); cps.push_back(c);}
  vgb_renderer* r=vgb_renderer_new(1,0,0); vgb_batch* b=vgb_batch_new(r);
  int reps=argc>1?atoi(argv[1]):20; size_t glyphs=0; 
  auto t0=std::chrono::steady_clock::now();
  for(int rep=0;rep<reps;++rep){ for(size_t i=0;i<fonts.size();++i){ vgb_batch_clear(b); for(uint32_t cp:cps[i]) if(cp<=0xFFFF) glyphs+=vgb_batch_add_glyph(b,fonts[i],cp);} }
  double dt=std::chrono::duration<double>(std::chrono::steady_clock::now()-t0).count();
  printf("%zu glyphs, %.3f us/glyph\n",glyphs,dt/glyphs*1e6);
}
