import numpy as np, chess_b200 as cb
toy=lambda n: 'tests/golden/toy/'+n
re_, rt, _ = cb.load_contacts(toy("ref.sparse"), toy("ref.bed"))
qe, qt, _ = cb.load_contacts(toy("qry.sparse"), toy("qry.bed"))
pairs,_ = cb.load_pairs_for_chroms(toy("pairs.bedpe"), set(rt)|set(qt))
eng = cb.CompareEngine(re_, rt, qe, qt)
a = eng.compare(pairs, kernel=0); b = eng.compare(pairs, kernel=1); c = eng.compare(pairs, kernel=2)
for k,(pid,r,q) in enumerate(pairs):
    print(pid, r, q, a[3][k], b[3][k], c[3][k], a[1][k], b[1][k], c[1][k], a[2][k], b[2][k])
