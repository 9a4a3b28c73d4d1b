import pickle, torch, sys
torch.manual_seed(0)
# distilled-data pickle as inference.py --load expects it: a list of (batch, 64, T) float tensors
data = [torch.randn(4, 64, 320) for _ in range(3)]
pickle.dump(data, open(sys.argv[1], "wb"))
