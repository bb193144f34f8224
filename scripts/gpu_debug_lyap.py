import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, warnings
warnings.simplefilter("ignore")
from scipy.stats import sem
from tests import scenarios
g = scenarios.golden("mg_lyap")
r = scenarios.mg_lyap(g, verification=False)
print("states maxrel", np.max(np.abs(r["states"]-g["states"])/np.abs(g["states"])))
print("weights sum", np.sum(r["weights"]), np.sum(g["weights"]), "accepted", r["accepted"], g["accepted"], "rejected", r["rejected"], g["rejected"])
for i in range(3):
    ours = np.average(r["lyaps"][:, i], weights=r["weights"]); theirs = np.average(g["lyaps"][:, i], weights=g["weights"])
    print(i, ours, theirs, abs(ours-theirs), 3.5*sem(g["lyaps"][:, i]))
