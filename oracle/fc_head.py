"""TEST INFRASTRUCTURE ONLY (tests/, __graft_entry__.smoke() and bench.py's CPU legs may import this; the product must not).

CPU restatement of the MC-CNN-accurate decision head. PARITY UNPINNED: the reference holds only the layer helper
`fc(input, num_in, num_out, name, relu)` = relu(xw_plus_b) with weights [num_in, num_out] (mc_cnn_brunch.py:95-106) and
never builds the head; the architecture is the "accurate" Middlebury net of the MC-CNN paper (Zbontar & LeCun, JMLR 2016:
3 hidden fully-connected layers of 384 units on the concatenated feature vectors, one sigmoid output, matching cost =
-similarity) on top of the tower the reference does have. Weight names follow the reference's variable scopes:
fc{1..4}/weights:0 ([128,384], [384,384], [384,384], [384,1]) and fc{1..4}/biases:0.
"""
import numpy as np

FC_UNITS = 384


def _relu(x):
    return np.maximum(x, 0)


def head_cost_volume(fl, fr, w, ndisp, fill=1.0, dtype=np.float32, emulate_fp16=False):
    """fl, fr: [H,W,64] features -> (CL, CR) [H,W,ndisp]: CL[y,x,d] = CR[y,x-d,d] = -sigmoid(net([fl[y,x]; fr[y,x-d]])),
    entries whose match falls outside the other image = fill (as the fast net's volume, process_functional.py:1111-1114).

    emulate_fp16=True rounds exactly where the CUDA kernel does (fc1 outputs, their sum, W2 / W3 and the fc2 output to
    fp16; everything else fp32), so the comparison isolates the kernel's data movement from its operand precision."""
    H, W, F = fl.shape
    t = np.float32 if emulate_fp16 else dtype
    W1 = w["fc1/weights:0"].astype(t)
    A1 = fl.astype(t) @ W1[:F] + w["fc1/biases:0"].astype(t)
    B1 = fr.astype(t) @ W1[F:]
    W2, W3 = w["fc2/weights:0"].astype(t), w["fc3/weights:0"].astype(t)
    if emulate_fp16:
        A1, B1 = A1.astype(np.float16), B1.astype(np.float16)
        W2, W3 = W2.astype(np.float16).astype(np.float32), W3.astype(np.float16).astype(np.float32)
    b2, b3 = w["fc2/biases:0"].astype(t), w["fc3/biases:0"].astype(t)
    w4, b4 = w["fc4/weights:0"].astype(t).reshape(-1), t(w["fc4/biases:0"].reshape(-1)[0])
    CL = np.full((H, W, ndisp), fill, np.float32)
    CR = np.full((H, W, ndisp), fill, np.float32)
    for d in range(min(ndisp, W)):
        h1 = _relu(A1[:, d:, :] + B1[:, :W - d, :])  # fp16 + fp16 -> fp16 (one rounding) when emulating
        h2 = _relu(h1.astype(t) @ W2 + b2)
        if emulate_fp16:
            h2 = h2.astype(np.float16).astype(np.float32)
        h3 = _relu(h2 @ W3 + b3)
        z = h3 @ w4 + b4
        cost = (-1.0 / (1.0 + np.exp(-z.astype(np.float64)))).astype(np.float32)
        CL[:, d:, d] = cost
        CR[:, :W - d, d] = cost
    return CL, CR
