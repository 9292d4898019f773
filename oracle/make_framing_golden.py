"""Golden packets from the UNMODIFIED reference framing code (test infrastructure, not product).

Runs ``/root/reference/src/neuralstego/codec/packet.py`` (``build_packet``) for a fixed ``msg_id`` and writes
``tests/golden/framing_packets.json``.  ``reedsolo`` is not installed in this image, so the reference itself can only
produce the ``ecc="none"`` packets; for ``ecc="rs"`` the reference's ``RSCodec`` symbol is pointed at this repo's
reedsolo-compatible codec (pinned separately by reedsolo's published known answer and the reference's own RS tests) and
the packet layout still comes from the reference's code.
"""
import base64, json, os, sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, "/root/reference/src")

import neuralstego.codec.packet as RP                      # the reference
from neuralsteganography_b200 import framing as F

MSG_ID = "00000000-1111-2222-3333-444444444444"
payloads = [b"", b"hello", bytes(range(256)), b"neural stego " * 40]
cases = []
for crc in (False, True):
    for ecc, nsym in (("none", 0), ("rs", 10), ("rs", 4)):
        if ecc == "rs":
            RP.RSCodec = F.RSCodec                          # see the docstring
            RP.ReedSolomonError = F.ReedSolomonError
        for seq, pl in enumerate(payloads):
            cfg = {"chunk_bytes": 256, "crc": crc, "ecc": ecc, "nsym": nsym}
            pkt = RP.build_packet(pl, msg_id=MSG_ID, seq=seq, total=len(payloads), cfg=cfg)
            back = RP.parse_packet(pkt, expected_cfg={"crc": crc, "ecc": ecc, "nsym": nsym})
            assert back.payload == pl
            cases.append({"cfg": cfg, "seq": seq, "total": len(payloads), "payload_b64": base64.b64encode(pl).decode(),
                          "packet": pkt.decode("utf-8"), "from": "reference" if ecc == "none" else "reference layout + repo RS"})
out = os.path.join(ROOT, "tests", "golden", "framing_packets.json")
json.dump({"msg_id": MSG_ID, "cases": cases}, open(out, "w"), indent=0)
print("wrote", out, len(cases), "cases")
