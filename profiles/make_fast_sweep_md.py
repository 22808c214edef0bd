"""profiles/fast_sweep.py output (jsonl) -> markdown table.   python profiles/make_fast_sweep_md.py sweep.jsonl > r01_fast_sweep.md"""
import json, sys
rows = [json.loads(l) for l in open(sys.argv[1]) if l.strip().startswith("{")]
summ = [r for r in rows if r.get("summary")]
rows = [r for r in rows if not r.get("summary")]
ROBUST = set("80bau3b adlittle afiro agg beaconfd blend boeing1 boeing2 bore3d cre-a cre-c czprob d6cube degen2 finnis fit1d fit1p fit2p ganges "
             "grow22 grow7 israel kb2 pilot87 pilotnov sc105 sc205 sc50a sc50b scfxm3 scorpion scrs8 scsd1 scsd8 sctap1 sctap2 sctap3 seba ship04l "
             "ship04s ship08l ship08s ship12l ship12s standata standgub standmps stocfor1 stocfor2 wood1p".split())
FRAGILE = set("25fv47 agg2 agg3 bandm bnl1 bnl2 brandy d2q06c degen3 e226 etamacro fffff800 forplan gfrd-pnc greenbea grow15 ken-07 ken-11 "
              "lotfi maros nesm pds-02 pds-06 pilot recipe scagr25 scagr7 scfxm1 scfxm2 scsd6 share1b share2b shell sierra woodw".split())
cls = lambda n: "R" if n in ROBUST else ("f" if n in FRAGILE else "-")
print("# Fast mode over the netlib fixtures (B200, device-resident hsd, `profiles/fast_sweep.py`)\n")
print("Per problem: status and printed iteration lines of fast mode against the reference's golden log, relative error of the final\n"
      "objective against the reference's, and whether all three north_star bars hold (same status, lines within 1, objective 1e-8).\n"
      "`class` is SURVEY.md H2's own classification of the REFERENCE under rounding-level perturbations (R = robust-50, f = fragile-35:\n"
      "recompiling the reference with FMA already changes its iteration count, - = not classified there).  Strict mode reproduces every one of these logs byte for byte.\n")
ok = sum(r["in_tol"] for r in rows)
print(f"**{ok} of {len(rows)} inside all tolerances; {sum(r['status'] == r['ref_status'] for r in rows)} end with the reference's status; "
      f"total solve time {sum(r['total_s'] for r in rows):.1f} s for the {len(rows)} problems.**\n")
print("| problem | class | N | status (ref) | lines (ref) | objective rel. err | in tolerance | total s | factor ms/call | solve ms/call |")
print("|---|---|---|---|---|---|---|---|---|---|")
for r in rows:
    print(f"| {r['name']} | {cls(r['name'])} | {r['N']} | {r['status']} ({r['ref_status']}) | {r['lines']} ({r['ref_lines']}) | "
          f"{r['obj_rel']:.1e} | {'yes' if r['in_tol'] else 'NO'} | {r['total_s']:.2f} | {r['factor_ms']:.2f} | {r['solve_ms']:.2f} |")
