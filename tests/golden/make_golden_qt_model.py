"""Decode the reference's only numeric linear fixture into a small JSON golden file.

Source: /root/reference/test/models_saved/linear_regressor_train_result.jls (Julia Serialization dump of
an MLJ machine).  The coefficient array `AB_t` is an Array{Float32}(6,4) whose serialized header
`15 00 0d 14 02 e5 e3` sits at byte 0xd6; the 96 data bytes follow (column-major).  The reference test
(test/computation_mpc_test.jl:1003-1006) uses AB = AB_t', A = AB[:, 1:4], B = AB[:, 5:end].

Run in the build container only (it reads /root/reference); the JSON it writes is committed.
"""
import json, pathlib
import numpy as np

SRC = pathlib.Path("/root/reference/test/models_saved/linear_regressor_train_result.jls")
OUT = pathlib.Path(__file__).with_name("qt_linear_model.json")

def main():
    raw = SRC.read_bytes()
    hdr = bytes([0x15, 0x00, 0x0D, 0x14, 0x02, 0xE5, 0xE3])
    off = raw.find(hdr)
    assert off == 0xD6 and raw.count(hdr) == 1, off
    ab_t = np.frombuffer(raw[off + 7: off + 7 + 96], dtype="<f4").reshape(4, 6)  # = AB (row-major view of AB_t')
    ab = ab_t.astype(np.float64)           # Float32 -> Float64 promotion is exact
    out = {
        "source": "test/models_saved/linear_regressor_train_result.jls @0xdd, Float32[6,4] column-major",
        "A": ab[:, :4].tolist(), "B": ab[:, 4:].tolist(),
        "A_f32_hex": [[np.float32(v).tobytes().hex() for v in row] for row in ab_t[:, :4]],
        "B_f32_hex": [[np.float32(v).tobytes().hex() for v in row] for row in ab_t[:, 4:]],
        # scenario of test/computation_mpc_test.jl:981-1040 and defaults src/main/main_mpc.jl:87-94
        "scenario": {"xmin": [0.2] * 4, "xmax": [1.36, 1.36, 1.30, 1.30], "umin": [0.0, 0.0], "umax": [4.0, 3.26],
                     "x_ref": [0.65] * 4, "u_ref": [1.2, 1.2], "x0": [0.6] * 4, "Q": 100.0, "R": 0.1, "S": 0.0},
    }
    OUT.write_text(json.dumps(out, indent=1))
    print("wrote", OUT)

if __name__ == "__main__":
    main()
