"""On-disk proof container of the reference's prover crate (prover/src/proof.rs:25-133, `Proof`): JSON object with base64
fields `proof`, `instances` (one instance column; every Fr as 32 bytes BIG-endian, i.e. `to_bytes()` reversed -- proof.rs:126-133,
read back at :77-85), `vk` (VerifyingKey bytes, SerdeFormat::Processed) and `git_version`.  Host-only marshalling: no field
arithmetic happens here (instance values are canonical integers; take them from the device with UOP_FROM_MONT)."""
import base64
import json

R_MOD = 21888242871839275222246405745257275088548364400416034343698204186575808495617


def serialize_instances(instances):
    """instances: list with exactly one column of canonical ints -> concatenated 32-byte big-endian words."""
    assert len(instances) == 1
    out = bytearray()
    for v in instances[0]:
        assert 0 <= v < R_MOD
        out += int(v).to_bytes(32, "big")
    return bytes(out)


def deserialize_instances(raw):
    assert len(raw) % 32 == 0
    return [[int.from_bytes(raw[i: i + 32], "big") for i in range(0, len(raw), 32)]]


class Proof:
    def __init__(self, proof, instances_raw, vk=b"", git_version=None):
        self.proof, self.instances_raw, self.vk, self.git_version = bytes(proof), bytes(instances_raw), bytes(vk), git_version

    @staticmethod
    def new(proof, instances, vk=b"", git_version=None):
        return Proof(proof, serialize_instances(instances), vk, git_version)

    def instances(self):
        return deserialize_instances(self.instances_raw)

    def to_json_obj(self):
        b64 = lambda b: base64.b64encode(b).decode()
        return {"proof": b64(self.proof), "instances": b64(self.instances_raw), "vk": b64(self.vk), "git_version": self.git_version}

    def to_json(self):
        return json.dumps(self.to_json_obj(), separators=(",", ":"))

    @staticmethod
    def from_json_obj(o):
        d = base64.b64decode
        return Proof(d(o["proof"]), d(o["instances"]), d(o["vk"]), o.get("git_version"))

    def dump(self, directory, filename):
        """full_proof_{filename}.json + vk_{filename}.vkey, as Proof::dump does (proof.rs:67-71,108-123)."""
        import os
        with open(os.path.join(directory, f"vk_{filename}.vkey"), "wb") as f:
            f.write(self.vk)
        with open(os.path.join(directory, f"full_proof_{filename}.json"), "w") as f:
            f.write(self.to_json())

    @staticmethod
    def from_json_file(directory, filename):
        import os
        with open(os.path.join(directory, f"full_proof_{filename}.json")) as f:
            return Proof.from_json_obj(json.load(f))
