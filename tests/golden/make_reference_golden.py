#!/usr/bin/env python
"""Generate golden vectors by RUNNING THE REFERENCE'S OWN FUNCTIONS (build container only).

The reference (mounted read-only at /root/reference) cannot be imported as-is here:
sqlalchemy / psycopg2 / python-jose / passlib are not installed and there is no
PostgreSQL.  This script installs *stub modules* for exactly those third-party
dependencies (an in-memory session with the handful of Query methods the hot-path
functions call), imports the reference's unmodified modules from /root/reference, runs

  app/services/webdav_sync.py        WebDAVSync._process_image_batch (+ _calculate_hash_from_bytes,
                                     _validate_image, _get_image_metadata, _download_and_process_image)
  app/services/activity_api_sync.py  ActivityAPISync._process_new_image
  app/api/routes/images.py           buscar_imagens_por_hash
  app/crud/classificacao_crud.py     obter_classificacoes_imagens, criar_ou_atualizar_classificacao
  app/api/routes/classificacoes.py   obter_contagem_classificacoes, listar_historico_usuario

on small seeded inputs, and writes inputs + outputs to tests/golden/reference_ingest.json
and tests/golden/reference_labels.json.  Those JSON files are committed; this script does
not run on the GPU box (nothing there may read /root/reference).  No reference source is
copied: the modules are imported from where they lie.

    python tests/golden/make_reference_golden.py
"""
from __future__ import annotations

import asyncio
import base64
import copy
import io
import json
import os
import sys
import types
import uuid
from datetime import datetime, timezone

HERE = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference"


# --------------------------------------------------------------------------- stub "ORM"
class IntegrityError(Exception):
    pass


class Pred:
    def __init__(self, fn):
        self.fn = fn


class Col:
    """Class-level column expression; instances shadow it with plain attributes."""

    def __set_name__(self, owner, name):
        self.owner, self.name = owner, name

    def __eq__(self, other):  # noqa: D105
        if isinstance(other, Col):
            return Pred(lambda r: True)  # join condition: ignored by the stub
        return Pred(lambda r, n=self.name, v=other: getattr(r, n) == v)

    def __hash__(self):
        return id(self)

    def in_(self, values):
        vals = set(values)
        return Pred(lambda r, n=self.name: getattr(r, n) in vals)


def make_model(name, cols, pk):
    def __init__(self, **kw):
        for c in cols:
            setattr(self, c, kw.get(c))
        for k, v in kw.items():
            setattr(self, k, v)

    ns = {c: Col() for c in cols}
    ns["__init__"] = __init__
    ns["__pk__"] = pk
    cls = type(name, (), ns)
    for c in cols:
        getattr(cls, c).__set_name__(cls, c)
    return cls


class Query:
    def __init__(self, session, entity):
        self.s = session
        if isinstance(entity, Col):
            self.project = entity.name
            self.rows = list(session.tables.setdefault(entity.owner, []))
        else:
            self.project = None
            self.rows = list(session.tables.setdefault(entity, []))

    def filter(self, *preds):
        for p in preds:
            self.rows = [r for r in self.rows if p.fn(r)]
        return self

    def filter_by(self, **kw):
        self.rows = [r for r in self.rows if all(getattr(r, k) == v for k, v in kw.items())]
        return self

    def _out(self):
        if self.project:
            return [(getattr(r, self.project),) for r in self.rows]
        return self.rows

    def distinct(self):
        seen, out = set(), []
        for r in self.rows:
            key = getattr(r, self.project) if self.project else id(r)
            if key not in seen:
                seen.add(key)
                out.append(r)
        self.rows = out
        return self

    def limit(self, n):
        self.rows = self.rows[:n]
        return self

    def all(self):
        return self._out()

    def first(self):
        out = self._out()
        return out[0] if out else None

    def count(self):
        return len(self.rows)

    def update(self, values, synchronize_session=False):
        for r in self.rows:
            for k, v in values.items():
                setattr(r, k, v)
        return len(self.rows)


class CannedQuery:
    """For the 5-way join in listar_historico_usuario: joins/filters are SQL-side, the
    stub returns the canned joined page so the Python grouping loop runs on it."""

    def __init__(self, rows):
        self.rows = rows

    def join(self, *a, **k):
        return self

    def filter(self, *a, **k):
        return self

    def order_by(self, *a, **k):
        return self

    def offset(self, n):
        return self

    def limit(self, n):
        return self

    def all(self):
        return self.rows

    def count(self):
        return len(self.rows)


class Session:
    def __init__(self):
        self.tables = {}
        self.pending = []
        self.canned = None

    def query(self, *entities):
        if len(entities) > 1:
            return CannedQuery(self.canned)
        return Query(self, entities[0])

    def add(self, obj):
        self.pending.append(obj)

    def bulk_save_objects(self, objs):
        for o in objs:
            self.pending.append(o)
        self.flush()

    def flush(self):
        for obj in self.pending:
            tbl = self.tables.setdefault(type(obj), [])
            pk = type(obj).__pk__
            if pk and any(getattr(r, pk) == getattr(obj, pk) for r in tbl):
                self.pending = []
                raise IntegrityError(f"duplicate key {pk}")
            tbl.append(obj)
        self.pending = []

    def commit(self):
        self.flush()

    def rollback(self):
        self.pending = []

    def refresh(self, obj):
        pass


def install_stubs():
    os.environ.setdefault("JWT_SECRET_KEY", "golden-generator")
    os.environ.setdefault("DATABASE_URL", "postgresql://stub")
    sa = types.ModuleType("sqlalchemy")
    sa.and_ = lambda *a: Pred(lambda r: all(p.fn(r) for p in a))
    sa.or_ = lambda *a: Pred(lambda r: any(p.fn(r) for p in a))
    sa.desc = lambda c: c
    sa_orm = types.ModuleType("sqlalchemy.orm")
    sa_orm.Session = Session
    sa_exc = types.ModuleType("sqlalchemy.exc")
    sa_exc.IntegrityError = IntegrityError
    sys.modules.update({"sqlalchemy": sa, "sqlalchemy.orm": sa_orm, "sqlalchemy.exc": sa_exc})

    models = types.ModuleType("app.db.models")
    models.Imagem = make_model(
        "Imagem",
        ["content_hash", "nome_img", "caminho_img", "metadados", "existe_no_nextcloud",
         "data_proc", "data_sinc", "id_cnj"], "content_hash")
    models.ConjuntoImagens = make_model(
        "ConjuntoImagens",
        ["id_cnj", "nome_conj", "caminho_conj", "file_id", "imagens_sincronizadas",
         "existe_no_nextcloud", "data_proc", "data_sinc", "id_amb"], "id_cnj")
    models.Classificacao = make_model(
        "Classificacao",
        ["id_cla", "data_criado", "data_modificado", "id_con", "id_img", "id_opc", "ativo"], None)
    models.Opcao = make_model("Opcao", ["id_opc", "texto", "id_amb"], "id_opc")
    models.Ambiente = make_model("Ambiente", ["id_amb", "titulo_amb"], "id_amb")
    models.UsuarioAmbienteProgresso = make_model(
        "UsuarioAmbienteProgresso",
        ["id_con", "id_amb", "ultimo_data_proc_processado", "ultimo_content_hash_processado",
         "total_classificadas", "data_ultima_atividade"], None)
    models.Usuario = make_model("Usuario", ["id_usu", "convencional"], "id_usu")
    models.AmbienteConjuntoImagens = make_model(
        "AmbienteConjuntoImagens", ["id_amb", "id_cnj", "ativo"], None)
    db_pkg = types.ModuleType("app.db")
    db_pkg.models = models
    db_pkg.__path__ = []
    database = types.ModuleType("app.db.database")
    database.get_db = lambda: None
    nc = types.ModuleType("app.services.nextcloud_service")
    nc.NextCloudClient = object
    auth = types.ModuleType("app.services.auth_service")
    auth.get_current_user = lambda: None
    sys.modules.update({
        "app.db": db_pkg, "app.db.models": models, "app.db.database": database,
        "app.services.nextcloud_service": nc, "app.services.auth_service": auth,
    })
    sys.path.insert(0, REF)
    return models


# --------------------------------------------------------------------------- fake NextCloud
class Resp:
    def __init__(self, content):
        self.content = content


class Client:
    base_url, auth, verify_ssl = "http://stub", ("u", "p"), False

    def __init__(self, files, failures):
        self.files, self.failures = files, failures

    def get_file(self, path):
        import requests
        kind = self.failures.get(path)
        if kind == "connection":
            raise requests.exceptions.ConnectionError("stub connection error")
        if kind == "timeout":
            raise requests.exceptions.Timeout("stub timeout")
        if kind == "other":
            raise ValueError("stub failure")
        return Resp(self.files[path])

    def list_folder(self, folder_path, depth=0):
        name = folder_path.rsplit("/", 1)[-1]
        return [{"is_collection": True, "file_id": "fid-" + (name or "root"), "name": name,
                 "path": folder_path}]


def png_bytes(w, h, color):
    from PIL import Image
    buf = io.BytesIO()
    Image.new("RGB", (w, h), color).save(buf, format="PNG")
    return buf.getvalue()


def jpeg_bytes(w, h, color):
    from PIL import Image
    buf = io.BytesIO()
    Image.new("L", (w, h), color).save(buf, format="JPEG")
    return buf.getvalue()


def b64(b):
    return base64.b64encode(b).decode("ascii")


def dump_imagens(session, models):
    out = {}
    for r in session.tables.get(models.Imagem, []):
        md = r.metadados or {}
        out[r.content_hash] = {
            "nome_img": r.nome_img,
            "caminho_img": r.caminho_img,
            "existe_no_nextcloud": r.existe_no_nextcloud,
            "id_cnj": str(r.id_cnj),
            "image_meta": copy.deepcopy(md.get("image")),          # snapshot: later batches mutate these dicts
            "nextcloud_meta": copy.deepcopy(md.get("nextcloud")),
            "sync_method": (md.get("sync") or {}).get("sync_method"),
            "first_seen": r.data_proc == r.data_sinc,
        }
    return out


def gen_ingest(models):
    import numpy as np
    from app.services.webdav_sync import WebDAVSync
    from app.services.activity_api_sync import ActivityAPISync
    from app.api.routes import images as images_route

    rng = np.random.default_rng(0xB200)
    raw = [rng.integers(0, 256, size=n, dtype=np.uint8).tobytes() for n in (0, 1, 55, 56, 63, 64, 65, 119, 120, 300, 4096)]
    png_a, png_b, jpg = png_bytes(5, 7, (1, 2, 3)), png_bytes(64, 48, (200, 10, 30)), jpeg_bytes(16, 9, 77)
    lm = datetime(2024, 5, 6, 7, 8, 9, tzinfo=timezone.utc)

    files, infos = {}, []

    def add(path, data, name=None, ctype="image/png", lm_=lm, **extra):
        files[path] = data
        info = {"path": path, "name": name if name is not None else path.rsplit("/", 1)[-1],
                "file_id": f"f{len(infos)}", "etag": f"e{len(infos)}", "content_type": ctype,
                "content_length": len(data), "last_modified": lm_}
        info.update(extra)
        infos.append(info)

    add("/set1/a.png", png_a)
    add("/set1/b.png", png_b)
    add("/set1/a_copy.png", png_a)                       # duplicate inside the batch
    add("/set1/raw0.jpg", raw[9], ctype="image/jpeg")    # headerless -> metadata {}
    add("/set1/notes.txt", raw[3], ctype="text/plain")   # bad extension
    add("/set1/fake.png", raw[4], ctype="application/octet-stream")  # bad MIME
    add("/set1/c.JPG", jpg, ctype="image/jpeg; charset=binary")
    add("/set1/down.png", raw[5])                        # connection error
    add("/set1/slow.png", raw[6])                        # timeout
    add("/set1/boom.png", raw[7])                        # other exception
    add("/set1/empty.gif", raw[0], ctype="image/gif")    # zero-length file still hashes
    add("/set1/a_third.webp", png_a, ctype="image/webp", lm_=None)
    for i, r in enumerate(raw[1:9]):
        add(f"/set1/r{i}.bmp", r, ctype="image/bmp")
    add("/set1/r2_again.tiff", raw[3], ctype="image/tiff")
    failures = {"/set1/down.png": "connection", "/set1/slow.png": "timeout", "/set1/boom.png": "other"}

    session = Session()
    client = Client(files, failures)
    sync = WebDAVSync(client, session)
    cid1, cid2 = uuid.UUID(int=1), uuid.UUID(int=2)

    batches = [infos[:8], infos[8:16], infos[16:], infos[:6]]   # last one re-syncs: all updates
    cids = [cid1, cid1, cid2, cid2]
    batch_out = []
    for b, cid in zip(batches, cids):
        stats = sync._process_image_batch(b, "/set1", cid)
        session.commit()
        batch_out.append({"indices": [infos.index(i) for i in b], "conjunto_id": str(cid),
                          "stats": stats, "table_after": dump_imagens(session, models)})

    singles = []
    for info in infos:
        h, meta = sync._download_and_process_image(info)
        singles.append({"valid": sync._validate_image(info), "hash": h, "metadata": meta})

    # Activity-API single-image variant on a fresh table
    session2 = Session()
    act = ActivityAPISync(client, session2)
    act_out = []
    for info in infos[:12] + infos[:3]:
        ok = act._process_new_image(info)
        act_out.append({"index": infos.index(info), "ok": ok})
    act_table = dump_imagens(session2, models)

    # upload lookup against the table left by the WebDAV batches
    class Upload:
        def __init__(self, filename, content_type, data):
            self.filename, self.content_type, self._d = filename, content_type, data

        async def read(self):
            return self._d

    uploads = [("a.png", "image/png", png_a), ("x.bin", "application/pdf", raw[9]),
               ("nobody.png", "image/png", raw[10]), ("none", None, png_b),
               ("c.jpg", "image/jpeg", jpg), ("raw.jpg", "image/jpeg", raw[9])]
    resp = asyncio.run(images_route.buscar_imagens_por_hash(
        files=[Upload(*u) for u in uploads], db=session))

    def ser(i):
        d = dict(i)
        d["last_modified"] = d["last_modified"].isoformat() if d["last_modified"] else None
        return d

    return {
        "_generated_by": "tests/golden/make_reference_golden.py (reference functions, stub session)",
        "files": {p: b64(d) for p, d in files.items()},
        "failures": failures,
        "infos": [ser(i) for i in infos],
        "webdav_batches": batch_out,
        "singles": singles,
        "activity": {"calls": act_out, "table_after": act_table},
        "upload_lookup": {
            "uploads": [{"filename": f, "content_type": c, "data": b64(d)} for f, c, d in uploads],
            "response": resp.model_dump(),
        },
    }


def gen_labels(models):
    import numpy as np
    from app.crud import classificacao_crud
    from app.api.routes import classificacoes as cls_route

    rng = np.random.default_rng(0xF1E155)
    users = [uuid.UUID(int=100 + i) for i in range(3)]
    amb = uuid.UUID(int=500)
    opts = [uuid.UUID(int=900 + i) for i in range(5)]
    hashes = [f"{i:064x}" for i in range(12)]
    session = Session()
    cnj = uuid.UUID(int=700)
    session.tables[models.ConjuntoImagens] = [models.ConjuntoImagens(id_cnj=cnj, id_amb=amb)]
    session.tables[models.AmbienteConjuntoImagens] = [
        models.AmbienteConjuntoImagens(id_amb=amb, id_cnj=cnj, ativo=True)]
    session.tables[models.Imagem] = [
        models.Imagem(content_hash=h, id_cnj=cnj, data_proc=datetime(2024, 1, 1, tzinfo=timezone.utc))
        for h in hashes]
    session.tables[models.Opcao] = [models.Opcao(id_opc=o, texto=f"op{i % 4}", id_amb=amb)
                                    for i, o in enumerate(opts)]   # op0 text appears twice
    rows = []
    for i in range(60):
        rows.append(models.Classificacao(
            id_cla=uuid.UUID(int=10_000 + i), id_con=users[int(rng.integers(0, 3))],
            id_img=hashes[int(rng.integers(0, 10))], id_opc=opts[int(rng.integers(0, 5))],
            ativo=bool(rng.random() < 0.7), data_criado=datetime(2024, 1, 1, tzinfo=timezone.utc)))
    session.tables[models.Classificacao] = rows

    def row_d(c):
        return {"id_cla": str(c.id_cla), "id_con": str(c.id_con), "id_img": c.id_img,
                "id_opc": str(c.id_opc), "ativo": c.ativo}

    table_before = [row_d(c) for c in rows]
    group_cases = []
    imgs = session.tables[models.Imagem]
    for u, sel in [(users[0], imgs[:12]), (users[1], imgs[2:7]), (users[2], []),
                   ("not-a-uuid", imgs[:3]), (str(users[0]), imgs[5:6])]:
        res = classificacao_crud.obter_classificacoes_imagens(session, str(u) if not isinstance(u, str) else u, sel)
        group_cases.append({
            "id_con": str(u), "images": [i.content_hash for i in sel],
            "result": {h: [str(c.id_cla) for c in lst] for h, lst in res.items()}})

    count_cases = []
    for u in users:
        usuario = models.Usuario(convencional=types.SimpleNamespace(id_con=u))
        count_cases.append({"id_con": str(u),
                            "total": cls_route.obter_contagem_classificacoes(usuario=usuario, db=session)["total"]})
    count_cases.append({"id_con": None, "total": cls_route.obter_contagem_classificacoes(
        usuario=models.Usuario(convencional=None), db=session)["total"]})

    # history grouping: canned joined page (classificacao, imagem, opcao, conjunto, ambiente)
    ambiente = models.Ambiente(id_amb=amb, titulo_amb="Ambiente A")
    opc_by_id = {o.id_opc: o for o in session.tables[models.Opcao]}
    img_by_hash = {i.content_hash: i for i in imgs}
    for i in imgs:
        i.nome_img, i.caminho_img = f"img_{i.content_hash[-2:]}.png", f"/set 1/img_{i.content_hash[-2:]}.png"
    page = [(c, img_by_hash[c.id_img], opc_by_id[c.id_opc], session.tables[models.ConjuntoImagens][0], ambiente)
            for c in rows if c.id_con == users[0] and c.ativo][:20]
    session.canned = page
    usuario = models.Usuario(convencional=types.SimpleNamespace(id_con=users[0]))
    hist = cls_route.listar_historico_usuario(id_amb=None, page=1, page_size=50, usuario=usuario, db=session)
    history = {
        "joined": [[c.id_img, o.texto, str(o.id_opc)] for c, _, o, _, _ in page],
        "items": [{"content_hash": it["content_hash"], "ids_opcoes": it["ids_opcoes"],
                   "opcao_escolhida": it["opcao_escolhida"], "url_img": it["url_img"]} for it in hist["items"]],
        "total": hist["total"],
    }

    # classification delta + counter rule
    delta_cases = []
    u = users[0]
    for h, wanted in [(hashes[10], [opts[0], opts[1]]), (hashes[10], [opts[1], opts[2]]),
                      (hashes[10], [opts[0]]), (hashes[11], [opts[3]]), (hashes[11], [opts[3]]),
                      (hashes[0], [opts[0], opts[4]])]:
        existing = [c for c in session.tables[models.Classificacao] if c.id_con == u and c.id_img == h]
        before_active = sorted(str(c.id_opc) for c in existing if c.ativo)
        before_inactive = sorted(str(c.id_opc) for c in existing if not c.ativo)
        prog = classificacao_crud.obter_progresso_usuario(session, str(u), str(amb))
        before_total = prog.total_classificadas
        res, novas = classificacao_crud.criar_ou_atualizar_classificacao(
            session, str(u), str(amb), h, [str(o) for o in wanted])
        existing = [c for c in session.tables[models.Classificacao] if c.id_con == u and c.id_img == h]
        delta_cases.append({
            "content_hash": h, "wanted": [str(o) for o in wanted],
            "before_active": before_active, "before_inactive": before_inactive,
            "after_active": sorted(str(c.id_opc) for c in existing if c.ativo),
            "after_inactive": sorted(str(c.id_opc) for c in existing if not c.ativo),
            "total_novas": novas, "result_opcs": sorted(str(c.id_opc) for c in res),
            "counter_delta": prog.total_classificadas - before_total})

    return {
        "_generated_by": "tests/golden/make_reference_golden.py (reference functions, stub session)",
        "classificacoes": table_before,
        "group_by_image": group_cases,
        "distinct_count": count_cases,
        "history": history,
        "delta": delta_cases,
    }


def main():
    models = install_stubs()
    import logging
    logging.disable(logging.CRITICAL)
    ingest = gen_ingest(models)
    labels = gen_labels(models)
    with open(os.path.join(HERE, "reference_ingest.json"), "w") as f:
        json.dump(ingest, f, indent=1, sort_keys=True)
    with open(os.path.join(HERE, "reference_labels.json"), "w") as f:
        json.dump(labels, f, indent=1, sort_keys=True)
    print("wrote reference_ingest.json, reference_labels.json")
    for b in ingest["webdav_batches"]:
        print(" webdav batch", b["stats"])
    print(" activity", [c["ok"] for c in ingest["activity"]["calls"]])
    print(" lookup", ingest["upload_lookup"]["response"]["total_encontradas"])
    print(" counts", labels["distinct_count"])
    print(" delta", [(d["total_novas"], d["counter_delta"]) for d in labels["delta"]])


if __name__ == "__main__":
    main()
